#!/usr/bin/env python
"""bench.py — particle-steps/s of the salva3d DFSPH step path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            native arm (this repo's CUDA engine)
  python bench.py --impl reference --gpus N --steps K ...   reference arm: the CPU restatement of salva's own
                                                            algorithm (oracle/, kind "port": the Rust reference
                                                            cannot be built here) on the box's host cores

One "step" = one LiquidWorld::step (liquid_world.rs:62) of the whole world.  N = 1 workload: BASELINE.json
configs[1] (1M-particle cube dam-break, DFSPH + XSPH viscosity).  Prints ONE JSON line on rank 0.
Timing: device time of every step from CUDA events recorded on the engine's own stream (sph_step_stats.step_ms),
W >= 3 warm-up steps, working set (particle state + neighbour lists) larger than L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/s (DFSPH 3D)"
UNIT = "particle-steps/s"

# SURVEY.md §8(d) algorithmic bytes per particle per launch (compulsory-traffic model)
BYTES = dict(grid=136, density_alpha=24, divergence_eval=40, divergence_update=52, fold=64, xsph=52, artificial=52,
             akinci=104, integrate=64, predict_density=44, pressure_update=52, positions=64)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.lines = []

    def mark(self):
        """Samples taken from here on belong to the timed region."""
        self.t_mark = len(self.lines)

    def start(self):
        self.t_mark = 0
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[self.t_mark:] if len(self.lines) - self.t_mark >= 2 else self.lines
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(name, n_override=None):
    from salva_b200 import scenes
    if name == "c1":
        return scenes.scene_c1()
    if name == "c2":
        return scenes.scene_c2(n_override or 100)
    if name == "c3":
        return scenes.scene_c3(n_override or 216)
    if name == "c4":
        return scenes.scene_c4() if not n_override else scenes.scene_c4(n_override, n_override // 2, n_override // 2)
    if name == "c5":
        return scenes.scene_c5(n_override or 100)
    raise SystemExit("unknown config " + name)


def scene_particles(sc):
    return sum(len(f["positions"]) for f in sc["fluids"]), sum(len(b["positions"]) for b in sc["boundaries"])


def force_kinds(sc):
    names = {0: "XSPHViscosity", 1: "ArtificialViscosity", 2: "Akinci2013SurfaceTension", 3: "Becker2009Elasticity"}
    return [names[k] for f in sc["fluids"] for k, _ in f.get("forces", [])]


def best_oracle_threads(sc):
    """The port is timed with whichever host thread count is FASTEST on this box (oversubscribed hyper-threads
    or cgroup-limited cores make `all threads` several times slower: profiles/r1_oracle_thread_scaling.json)."""
    from oracle.oracle import OracleWorld
    from salva_b200 import scenes
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (4, 8, 16, 32, 64, ncpu) if t <= ncpu})
    best, best_t = None, None
    for th in cands:
        w = OracleWorld(sc["particle_radius"], sc["smoothing_factor"], solver=sc["solver"], sort_contacts=False,
                        num_threads=th)
        scenes.populate(w, sc)
        w.step(sc["dt"], sc["gravity"])
        t0 = time.perf_counter()
        w.step(sc["dt"], sc["gravity"])
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
    return best


def run_oracle(sc, steps, warmup, threads=0):
    """CPU restatement of the reference algorithm (oracle/, test infrastructure) timed on host cores."""
    from oracle.oracle import OracleWorld
    from salva_b200 import scenes
    if threads <= 0:
        threads = best_oracle_threads(sc)
    w = OracleWorld(sc["particle_radius"], sc["smoothing_factor"], solver=sc["solver"], sort_contacts=False,
                    num_threads=threads)
    scenes.populate(w, sc)
    for _ in range(warmup):
        w.step(sc["dt"], sc["gravity"])
    t0 = time.perf_counter()
    iters = []
    for _ in range(steps):
        w.step(sc["dt"], sc["gravity"])
        st = w.stats()
        iters.append((st["n_divergence_iter"], st["n_pressure_iter"]))
    dt = time.perf_counter() - t0
    nf, _ = scene_particles(sc)
    return nf * steps / dt, dt / steps * 1e3, w.stats()["threads"], iters


def reference_arm(args, rank):
    """--impl reference: the reference's own CPU algorithm on the box's host cores (oracle port)."""
    if rank != 0:
        return
    n = args.ref_n
    sc = build_scene(args.config, n)
    nf, nb = scene_particles(sc)
    value, ms, threads, iters = run_oracle(sc, args.steps, args.warmup)
    full = build_scene_name(args.config, args.n)
    sample = "%s scene at %d fluid particles (same generator, best of 4..nproc host threads = %d), %d+%d steps" % (
        args.config.upper(), nf, threads, args.warmup, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": full, "sample_particles": nf, "boundary_particles": nb, "solver": "DFSPH",
                       "forces": force_kinds(sc), "iterations_last_step": list(iters[-1]) if iters else None},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def build_scene_name(cfg, n):
    desc = {"c1": "C1 examples3d/basic3.rs 3375 particles", "c2": "C2 1M-particle cube dam-break, DFSPH + XSPH viscosity",
            "c3": "C3 10M particles DFSPH + Akinci2013 surface tension", "c4": "C4 32M particles DFSPH slab split",
            "c5": "C5 2M particles IISPH 2 fluids"}[cfg]
    return desc + (" (lattice edge %d)" % n if n else "")


def native_arm(args, rank, world_size):
    import torch
    import torch.distributed as dist
    from salva_b200 import DFSPHSolver, IISPHSolver, LiquidWorld, scenes

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world_size > 1:
        # weak scaling: the C2 block is repeated along x (slab axis), 1M particles per GPU, one shared tank;
        # ranks own x-slabs and exchange one-cell ghost columns through NCCL every Jacobi sub-iteration.
        from salva_b200 import slab
        from salva_b200.liquid_world import nccl_unique_id
        n = args.n or 100
        sc = scenes._dam_break(n * world_size, n, n, 0.025, 1.0 / 1000.0, [scenes.xsph_viscosity(0.5, 0.0)],
                               name="C2x%d-slabs" % world_size, tank_x_factor=1.0 + 1.0 / world_size)
    else:
        sc = build_scene(args.config, args.n)
    nf, nb = scene_particles(sc)
    solver = DFSPHSolver() if sc["solver"] == 0 else IISPHSolver()
    world = LiquidWorld(solver, particle_radius=sc["particle_radius"], smoothing_factor=sc["smoothing_factor"],
                        device=local_rank, deterministic=not args.fast_sort, gather_backend=args.backend)
    if world_size > 1:
        uid = slab.broadcast_unique_id(nccl_unique_id, rank, device=torch.device("cuda", local_rank))
        fh, _ = slab.populate_slab(world, sc, rank, world_size, uid)
    else:
        fh, _ = scenes.populate(world, sc)
    if args.force_iters:
        world.force_iterations(*args.force_iters)
    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs ~0.2 s to emit its first sample: start it before the warm-up steps
    for _ in range(warm):
        world.step(sc["dt"], sc["gravity"])

    # ---- timed region: device-resident inputs, CUDA-event time of every step -------------------------
    barrier()
    sampler.mark()
    t0 = time.perf_counter()
    acc = {}
    launches = 0
    iters = []
    for _ in range(args.steps):
        world.step(sc["dt"], sc["gravity"])
        st = world.stats()
        for k, v in st.items():
            if k.endswith("_ms"):
                acc[k] = acc.get(k, 0.0) + v
        acc["n_pressure_eval"] = acc.get("n_pressure_eval", 0) + st["n_pressure_eval"]
        acc["n_pressure_iter"] = acc.get("n_pressure_iter", 0) + st["n_pressure_iter"]
        acc["n_divergence_eval"] = acc.get("n_divergence_eval", 0) + st["n_divergence_eval"]
        acc["n_divergence_iter"] = acc.get("n_divergence_iter", 0) + st["n_divergence_iter"]
        launches += st["kernel_launches"]
        iters.append((st["n_divergence_iter"], st["n_pressure_iter"]))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_s = acc["step_ms"] * 1e-3
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall = float(t[0]), float(t[1])
    total_particles = nf  # all ranks together step the ONE world of nf particles (nf grows with N: weak scaling)
    value = total_particles * args.steps / dev_s

    # ---- e2e: the reference-facing call sequence with HOST buffers, copies inside the timed region ----
    f0 = fh[0]
    n0 = world.num_particles(f0)
    cap0 = int(n0 * 1.25) + 1024  # slab worlds gain / lose particles through migration
    hp_all = torch.empty((cap0, 3), dtype=torch.float32, pin_memory=True).numpy()
    hv_all = torch.empty((cap0, 3), dtype=torch.float32, pin_memory=True).numpy()
    world.read_fluid(f0, hp_all[:n0], hv_all[:n0])
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        world.write_fluid(f0, hp_all[:n0], hv_all[:n0])   # host edits of fluid.positions / velocities go in
        world.step(sc["dt"], sc["gravity"])               # LiquidWorld::step
        n0 = world.num_particles(f0)
        world.read_fluid(f0, hp_all[:n0], hv_all[:n0])    # results come back in original index order
    barrier()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_particles * e2e_steps / float(t[0])

    if rank != 0:
        return
    peak, peak_src = peaks()
    # ---- roofline of the pressure iteration kernels (K8a + K8b), algorithmic bytes / CUDA-event time ----
    n_eval, n_upd = acc["n_pressure_eval"], acc["n_pressure_iter"]
    pa_ms = acc["predict_density_ms"] / max(n_eval, 1)
    pb_ms = acc["pressure_update_ms"] / max(n_upd, 1)
    it_bytes = nf * (BYTES["predict_density"] + BYTES["pressure_update"])
    it_ms = pa_ms + pb_ms
    achieved = it_bytes / (it_ms * 1e-3) / 1e9 if it_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = (json.load(open(tp)).get(args.config) or {}).get("pair")
    roofline = {"bound": "hbm", "kernel": "k_predict_density + k_pressure_update (one DFSPH pressure iteration)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_launch": it_bytes, "ms_per_launch_pair": it_ms,
                "predict_density_ms": pa_ms, "pressure_update_ms": pb_ms,
                "limiter": "L1/LSU gather + FP32 issue, not DRAM (see DESIGN.md)"}

    # ---- cpu_baseline: oracle port on a bounded sample of the same workload ------------------------------
    cpu = None
    if not args.no_cpu:
        csc = build_scene(args.config, args.ref_n)
        cnf, _ = scene_particles(csc)
        cv, cms, threads, _ = run_oracle(csc, args.cpu_steps, 1)
        cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%s scene at %d fluid particles, 1+%d steps, %.0f ms/step" % (args.config.upper(), cnf,
                                                                                     args.cpu_steps, cms)}
    phases = {k: acc[k] / args.steps for k in sorted(acc) if k.endswith("_ms")}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": warm,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": build_scene_name(args.config, args.n), "fluid_particles_total": nf, "fluid_particles_per_gpu": nf // world_size,
                       "boundary_particles": nb, "solver": "DFSPH" if sc["solver"] == 0 else "IISPH",
                       "forces": force_kinds(sc), "dt": sc["dt"], "particle_radius": sc["particle_radius"],
                       "parallelism": "1 GPU" if world_size == 1 else
                       "%d x-slabs, 1-cell ghost columns via ncclSend/Recv per sub-iteration, %d exchanges/step" % (world_size, st["n_exchanges"]),
                       "iterations_per_step_mean": [float(np.mean([i[0] for i in iters])),
                                                    float(np.mean([i[1] for i in iters]))],
                       "l2": "working set (state + neighbour lists, ~%.0f MB) exceeds the 126 MB L2" %
                             ((nf * (16 * 6 + 4 * 8 + 64 * 4)) / 1e6),
                       "phase_ms_per_step": phases, "wall_ms_per_step": wall / args.steps * 1e3},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n0 * 24), "d2h_bytes_per_step": int(n0 * 24),
                    "steps": e2e_steps, "api": "sph_fluid_write + sph_world_step + sph_fluid_read (pinned host buffers)"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--n", type=int, default=0, help="override lattice edge (testing)")
    ap.add_argument("--ref-n", type=int, default=64, help="lattice edge of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fast-sort", action="store_true", help="skip the deterministic in-cell ordering")
    ap.add_argument("--force-iters", type=int, nargs=2, default=None)
    ap.add_argument("--backend", type=int, default=0, help="0 = L1/texture gathers (default), 1 = tile/TMA shared-memory gathers")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world_size = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    native_arm(args, rank, world_size)


if __name__ == "__main__":
    main()

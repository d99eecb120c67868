"""torchrun worker for tests/test_gpu_slab.py: N ranks step their slabs of one scene (NCCL ghost exchange + migration);
rank 0 also steps the same scene on one GPU and compares by particle id."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from salva_b200 import LiquidWorld, scenes, slab  # noqa: E402
from salva_b200.liquid_world import nccl_unique_id  # noqa: E402


def make_scene(kind):
    r = 0.05
    rng = np.random.default_rng(41)
    nx, ny, nz = 20, 8, 7
    pts = scenes.jitter(scenes.block_lattice(nx, ny, nz, r * 0.95), r, 43, amplitude=0.3)
    vel = rng.normal(0, 0.3, pts.shape).astype(np.float32)
    vel[:, 0] += np.where(pts[:, 0] < pts[:, 0].mean(), 1.5, -1.5).astype(np.float32)  # push particles across the planes
    tank = scenes.open_tank((-r, -r, -r), (nx * 2 * r + r, 1.2, nz * 2 * r + r), r)
    forces = {"xsph": [scenes.xsph_viscosity(0.5, 0.2)], "akinci": [scenes.akinci2013_surface_tension(1.0, 0.3)],
              "artificial": [scenes.artificial_viscosity(1.0, 0.0)], "he2014": [scenes.he2014_surface_tension(40.0, 30.0)],
              "wcsph": [scenes.wcsph_surface_tension(2.0)]}[kind]
    return dict(particle_radius=r, smoothing_factor=2.0, dt=0.004,
                fluids=[dict(positions=pts, velocities=vel, density0=1000.0, forces=forces)], boundaries=[dict(positions=tank)])


def main():
    kind, steps, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    free_running = len(sys.argv) > 4 and sys.argv[4] == "free"
    do_rebalance = len(sys.argv) > 4 and sys.argv[4] == "rebalance"
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, ws = dist.get_rank(), dist.get_world_size()
    sc = make_scene(kind)
    uid = slab.broadcast_unique_id(nccl_unique_id, rank, device=torch.device("cuda", local_rank))
    w = LiquidWorld(particle_radius=sc["particle_radius"], device=local_rank)
    planes = None
    if do_rebalance:  # a deliberately bad split: rank 0 starts with a quarter of the cell columns
        h = np.float32(sc["particle_radius"]) * np.float32(4.0)
        cols = slab.cell_columns(sc["fluids"][0]["positions"], h)
        lo, hi = int(cols.min()), int(cols.max()) + 1
        step = max(2, (hi - lo) // (2 * ws))
        planes = [slab.INT32_MIN] + [lo + step * r for r in range(1, ws)] + [slab.INT32_MAX]
    fh, _ = slab.populate_slab(w, sc, rank, ws, uid, planes)
    if not free_running:
        w.force_iterations(2, 3)
    n0 = w.num_particles(fh[0])
    migrated = 0
    iters = []
    rebalanced, imb_before, imb_after = 0, 0.0, 0.0
    for k in range(steps):
        if do_rebalance and k == steps // 2:
            imb_before, did = slab.rebalance(w, fh[0], threshold=0.05)
            rebalanced += int(did)
            imb_after, _ = slab.imbalance(w.num_particles(fh[0]))
        w.step(sc["dt"])
        st = w.stats()
        migrated += st["n_migrated"]
        iters.append((st["n_divergence_iter"], st["n_pressure_iter"]))
    p, v = w.read_fluid(fh[0])
    ids = w.read_ids(fh[0])
    gathered = [None] * ws
    dist.gather_object(dict(ids=ids, p=p, v=v, n0=n0, migrated=migrated, ghosts=st["n_ghost_particles"], exchanges=st["n_exchanges"]),
                       gathered if rank == 0 else None, dst=0)
    if rank == 0:
        ids = np.concatenate([g["ids"] for g in gathered])
        p = np.concatenate([g["p"] for g in gathered])
        v = np.concatenate([g["v"] for g in gathered])
        order = np.argsort(ids)
        ref = LiquidWorld(particle_radius=sc["particle_radius"], device=local_rank)
        fr, _ = scenes.populate(ref, sc)
        if not free_running:
            ref.force_iterations(2, 3)
        ref_iters = []
        for _ in range(steps):
            ref.step(sc["dt"])
            s2 = ref.stats()
            ref_iters.append((s2["n_divergence_iter"], s2["n_pressure_iter"]))
        pr, vr = ref.read_fluid(fr[0])
        h = float(ref.h)
        res = dict(n_total=int(len(ids)), n_expected=int(len(pr)), ids_unique=bool(len(np.unique(ids)) == len(ids)),
                   max_dx_over_h=float(np.abs(p[order] - pr).max() / h), max_dv=float(np.abs(v[order] - vr).max()),
                   h_over_dt=h / sc["dt"], migrated=int(sum(g["migrated"] for g in gathered)),
                   ghosts=[int(g["ghosts"]) for g in gathered], exchanges=[int(g["exchanges"]) for g in gathered],
                   n_per_rank=[int(len(g["ids"])) for g in gathered], n0_per_rank=[int(g["n0"]) for g in gathered],
                   iters_match=iters == ref_iters, iters=iters[-1], ref_iters=ref_iters[-1], rebalanced=rebalanced,
                   imbalance_before=imb_before, imbalance_after=imb_after)
        json.dump(res, open(out, "w"))
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

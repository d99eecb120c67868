#!/usr/bin/env python
"""GPU experiment driver (run under gpurun): times bench.py for a list of (library build, env toggles) variants and
prints one compact line per variant.  Usage: python tools/exp_variants.py <config> <steps> name=lib[,ENV=V...] ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    cfg, steps = sys.argv[1], sys.argv[2]
    out = []
    for spec in sys.argv[3:]:
        name, rest = spec.split("=", 1)
        parts = rest.split(",")
        env = dict(os.environ)
        env["SALVA_B200_LIB"] = os.path.join(ROOT, parts[0])
        for kv in parts[1:]:
            k, v = kv.split("=")
            env[k] = v
        extra = env.pop("BENCH_ARGS", "").split()
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", cfg, "--steps", steps, "--warmup", "3", "--no-cpu",
                            "--no-parity", "--no-settled"] + extra, env=env, capture_output=True, text=True)
        line = None
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                line = json.loads(ln)
        if line is None:
            print("%-28s FAILED rc=%d %s" % (name, r.returncode, (r.stderr or "")[-300:].replace("\n", " | ")), flush=True)
            continue
        ph = line.get("phases") or line["config"].get("phase_ms_per_step") or {}
        rf = line["roofline"]
        msg = ("%-28s step %.4f ms | pred %.4f upd %.4f pair %.4f frac %.3f | grid %.3f nbr %.3f dens %.3f div %.3f force %.3f press %.3f integ %.3f | it %s | wall %.3f"
               % (name, line["ms_per_step"], rf["predict_density_ms"], rf["pressure_update_ms"], rf["ms_per_launch_pair"], rf["frac"],
                  ph.get("grid_ms", 0), ph.get("neighbors_ms", 0), ph.get("density_ms", 0), ph.get("divergence_ms", 0), ph.get("nonpressure_ms", 0),
                  ph.get("pressure_ms", 0), ph.get("integrate_ms", 0), line.get("iterations_per_step_mean") or line["config"].get("iterations_per_step_mean"),
                  line.get("wall_ms_per_step") or line["config"].get("wall_ms_per_step", 0)))
        print(msg, flush=True)
        out.append({"name": name, "line": line})
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "exp_%s.json" % cfg), "w"))


if __name__ == "__main__":
    main()

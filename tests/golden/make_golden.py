"""Generates tests/golden/*.npz from the CPU oracle (oracle/oracle.cpp).

The reference itself cannot run here (Rust, no toolchain) and ships no numeric fixtures for this path, so these vectors
are ORACLE outputs (PARITY UNPINNED upstream): they freeze the oracle's behaviour so that (a) oracle regressions and
(b) GPU drift are both caught against the same committed numbers.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleWorld  # noqa: E402
from salva_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(scene, solver, steps, n_div, n_press):
    w = OracleWorld(scene["particle_radius"], scene["smoothing_factor"], solver=solver, num_threads=4)
    fh, bh = scenes.populate(w, scene)
    w.force_iterations(n_div, n_press)
    for _ in range(steps):
        w.step(scene["dt"], scene["gravity"])
    out = {}
    for k, h in enumerate(fh):
        p, v = w.read_fluid(h)
        out["pos%d" % k], out["vel%d" % k] = p, v
        out["density%d" % k] = w.debug(h, "density")
        out["ncontacts%d" % k] = w.debug(h, "num_fluid_contacts").astype(np.int32)
        out["vc%d" % k] = w.debug(h, "velocity_change")
    return out


def main():
    c1 = scenes.scene_c1()
    np.savez_compressed(os.path.join(HERE, "c1_basic3_dfsph_5steps.npz"), **run(c1, 0, 5, 1, 2))
    c5 = scenes.scene_c5(6)
    np.savez_compressed(os.path.join(HERE, "c5_small_iisph_4steps.npz"), **run(c5, 1, 4, -1, 3))
    ts = scenes.scene_tension_small()
    np.savez_compressed(os.path.join(HERE, "tension_small_dfsph_5steps.npz"), **run(ts, 0, 5, 2, 3))
    print("written")


if __name__ == "__main__":
    main()

"""Oracle thread-scaling probe (which host thread count is the fairest CPU baseline on this box)."""
import sys, time, json
sys.path.insert(0, ".")
from oracle.oracle import OracleWorld
from salva_b200 import scenes
sc = scenes.scene_c2(64)
out = {}
for th in (1, 8, 16, 32, 64, 128):
    w = OracleWorld(sc["particle_radius"], 2.0, sort_contacts=False, num_threads=th)
    scenes.populate(w, sc)
    w.step(sc["dt"])
    t0 = time.perf_counter()
    for _ in range(2):
        w.step(sc["dt"])
    dt = (time.perf_counter() - t0) / 2
    st = w.stats()
    out[th] = dict(ms=dt * 1e3, grid=st["grid_ms"], nbr=st["neighbors_ms"], dens=st["density_ms"], div=st["divergence_ms"], press=st["pressure_ms"])
    print(th, out[th], flush=True)
json.dump(out, open("gpurun_out/oracle_threads.json", "w"))

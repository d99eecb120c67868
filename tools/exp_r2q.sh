#!/bin/bash
# (not run in round 2: written after the GPU budget was spent)  Row order (SALVA_B200_XYSUB=2) on one GPU: the whole GPU test suite in
# that mode, then bench.py with explicit orders on C3 / C2 / C4-slice, then one ncu capture of the pair for the sector / wavefront counts
# of profiles/r2_l1tex_wavefront_model.md.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(SALVA_B200_XYSUB=2 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/pytest_gpu_rows.txt
for cfg in c3 c2 c4; do
  for order in h rows; do
    timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --grid-order $order > $O/bench_${cfg}_$order.json 2> $O/bench_${cfg}_$order.err
  done
done
SALVA_B200_XYSUB=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vel_(divergence|update)_u|k_neighbors" -s 8 -c 4 -f \
    -o $O/r2q_rows_c3 python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-settled --grid-order rows > $O/ncu_r2q.log 2>&1
python - <<PY
import json
for cfg in ("c3", "c2", "c4"):
    for order in ("h", "rows"):
        try:
            d = json.loads([l for l in open("$O/bench_%s_%s.json" % (cfg, order)) if l.startswith("{")][-1])
            print(cfg, order, "ms/step %.4f pair %.4f nbr %.3f grid %.3f parity %s settled %s" % (
                d["ms_per_step"], d["roofline"]["ms_per_launch_pair"], d["phases"]["neighbors_ms"], d["phases"]["grid_ms"], d["parity"]["ok"],
                (d.get("settled") or {}).get("ms_per_step")))
        except Exception as e:
            print(cfg, order, "FAILED", e)
PY
cat $O/pytest_gpu_rows.txt

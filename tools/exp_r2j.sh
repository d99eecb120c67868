#!/bin/bash
# round-2 final records of the committed build (one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/pytest_gpu_r2j.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2j.txt 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_c3_r2j.json 2> $O/bench_c3_r2j.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > $O/bench_c2_r2j.json 2> $O/bench_c2_r2j.err
timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu > $O/bench_c4slice_r2j.json 2> $O/bench_c4slice_r2j.err
tail -2 $O/pytest_gpu_r2j.txt; tail -1 $O/smoke_r2j.txt; python - <<PY
import json
for f in ("bench_c3_r2j", "bench_c2_r2j", "bench_c4slice_r2j"):
    try:
        d = json.loads([l for l in open("$O/%s.json" % f) if l.startswith("{")][-1])
        s = d.get("settled") or {}
        print(f, "value %.4g ms/step %.4f frac %.4f parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], (d.get("parity") or {}).get("ok")),
              "settled", s.get("ms_per_step"), s.get("iterations_per_step_mean"), s.get("grid_dims"), (s.get("phases") or {}).get("grid_ms"))
    except Exception as e:
        print(f, "FAILED", e)
PY

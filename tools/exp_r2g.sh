#!/bin/bash
# round-2 last check of the committed build on one GPU: GPU tests, smoke, the driver's two bench invocations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $O/pytest_gpu_r2g.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2g.txt 2>&1
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference_r2g.json 2> $O/bench_reference_r2g.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_c3_r2g.json 2> $O/bench_c3_r2g.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > $O/bench_c2_r2g.json 2> $O/bench_c2_r2g.err
tail -2 $O/pytest_gpu_r2g.txt; tail -1 $O/smoke_r2g.txt; python - <<PY
import json
for f in ("bench_c3_r2g", "bench_c2_r2g", "bench_reference_r2g"):
    try:
        d = json.loads([l for l in open("$O/%s.json" % f) if l.startswith("{")][-1])
        print(f, "value %.4g ms/step %.4f" % (d["value"], d["ms_per_step"]), (d.get("roofline") or {}).get("frac"), (d.get("parity") or {}).get("ok"), d.get("phases"))
    except Exception as e:
        print(f, "FAILED", e)
PY

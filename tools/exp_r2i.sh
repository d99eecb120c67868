#!/bin/bash
# round-2 experiment I: fused reorder + v*, fused fold + integrate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/pytest_gpu_r2i.txt
L=salva_b200/libsalva_b200.so
echo "== variants" > $O/exp_r2i.txt
for cfg in c2 c3 c4; do
  echo "== $cfg" >> $O/exp_r2i.txt
  timeout 900 python tools/exp_variants.py $cfg 10 default=$L nofuse_fold=$L,SALVA_B200_FUSE_FOLD=0 >> $O/exp_r2i.txt 2>&1
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_c3_r2i.json 2> $O/bench_c3_r2i.err
cat $O/exp_r2i.txt; tail -2 $O/pytest_gpu_r2i.txt; python - <<PY
import json
d = json.loads([l for l in open("$O/bench_c3_r2i.json") if l.startswith("{")][-1])
print("c3", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity"]["ok"], d["settled"])
PY

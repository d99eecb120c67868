#!/bin/bash
# round-2 experiment M (one GPU): packed-f32 (FADD2 / FMUL2) candidate test in the neighbour search — all GPU tests, A/B on C3 / C2 / C4 slice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/pytest_gpu_r2m.txt
L=salva_b200/libsalva_b200.so
rm -f $O/exp_r2m.txt
for cfg in c3 c2 c4; do
  echo "== $cfg" >> $O/exp_r2m.txt
  timeout 600 python tools/exp_variants.py $cfg 10 packed=$L scalar=salva_b200/variants/v_nopacked.so packed_again=$L >> $O/exp_r2m.txt 2>&1
done
cat $O/exp_r2m.txt; cat $O/pytest_gpu_r2m.txt

#!/bin/bash
# round-2 experiment H: neighbour-search candidates split over TEX / LSU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(SALVA_B200_NBR_TEX=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -4) > $O/pytest_gpu_r2h_nbrtex.txt
L=salva_b200/libsalva_b200.so
echo "== variants" > $O/exp_r2h.txt
for cfg in c2 c3; do
  echo "== $cfg" >> $O/exp_r2h.txt
  timeout 900 python tools/exp_variants.py $cfg 10 default=$L nbr_tex=$L,SALVA_B200_NBR_TEX=1 >> $O/exp_r2h.txt 2>&1
done
cat $O/exp_r2h.txt; cat $O/pytest_gpu_r2h_nbrtex.txt

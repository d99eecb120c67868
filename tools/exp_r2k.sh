#!/bin/bash
# round-2 experiment K (one GPU): z-binned counting sort (Consts::zsub) - all GPU tests incl. the full-size ones, then the
# neighbour-search cost for zsub = 1 (plain h-cells), 2, 3, 4 on C3 and C2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16) > $O/pytest_gpu_r2k.txt
L=salva_b200/libsalva_b200.so
for cfg in c3 c2 c4; do
  echo "== $cfg" >> $O/exp_r2k.txt
  timeout 600 python tools/exp_variants.py $cfg 10 zsub1=$L,SALVA_B200_ZSUB=1 zsub2=$L,SALVA_B200_ZSUB=2 zsub3=$L,SALVA_B200_ZSUB=3 zsub4=$L,SALVA_B200_ZSUB=4 \
      >> $O/exp_r2k.txt 2>&1
done
cat $O/exp_r2k.txt; cat $O/pytest_gpu_r2k.txt

#!/bin/bash
# round-2 experiment L (one GPU): running-pointer list stores in k_neighbors (A/B), quick parity subset, and one
# `ncu --set full` capture of k_neighbors at C3 to find what bounds it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4) > $O/pytest_gpu_r2l.txt
L=salva_b200/libsalva_b200.so
rm -f $O/exp_r2l.txt
for cfg in c3 c2; do
  echo "== $cfg" >> $O/exp_r2l.txt
  timeout 600 python tools/exp_variants.py $cfg 10 runptr=$L norunptr=salva_b200/variants/v_norunptr.so runptr_again=$L >> $O/exp_r2l.txt 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_neighbors -s 3 -c 1 -f -o $O/r2l_nbr_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-settled > $O/ncu_r2l.log 2>&1
ls -la $O/r2l_nbr_c3.ncu-rep
cat $O/exp_r2l.txt; cat $O/pytest_gpu_r2l.txt

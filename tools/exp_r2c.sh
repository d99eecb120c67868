#!/bin/bash
# round-2 experiment C (1 GPU): pipe-split defaults, bounds carried over from the previous step, full bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > $O/pytest_gpu_r2c.txt
L=salva_b200/libsalva_b200.so
echo "== variants" > $O/exp_r2c.txt
for cfg in c2 c3; do
  echo "== $cfg" >> $O/exp_r2c.txt
  timeout 1200 python tools/exp_variants.py $cfg 10 \
    default=$L \
    eval_old=$L,SALVA_B200_UNI_EVAL=2 \
    upd_lsu=$L,SALVA_B200_UNI_UPD=2 \
    upd_alt_kernel=$L,SALVA_B200_UNI_UPD=3 \
    >> $O/exp_r2c.txt 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3_default.json 2> $O/bench_c3_default.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --config c1 --steps 50 --warmup 5 --no-cpu --no-parity --no-settled > $O/bench_c1.json 2> $O/bench_c1.err
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu --no-parity --no-settled > $O/bench_c5.json 2> $O/bench_c5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vel_(divergence|update)_u" -s 8 -c 3 -o $O/r2c_pair_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-settled > $O/ncu_full_c3.log 2>&1
cat $O/exp_r2c.txt; tail -3 $O/pytest_gpu_r2c.txt; tail -c 600 $O/bench_c3_default.err

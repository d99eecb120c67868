#!/bin/bash
# round-2 final 1-GPU validation: GPU test suite, coalesced list stores, persistent-CTA experiment, full bench lines, ncu of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > $O/pytest_gpu_r2f.txt
L=salva_b200/libsalva_b200.so
echo "== variants" > $O/exp_r2f.txt
for cfg in c2 c3; do
  echo "== $cfg" >> $O/exp_r2f.txt
  timeout 900 python tools/exp_variants.py $cfg 10 default=$L upd_persist=$L,SALVA_B200_UNI_UPD=4 >> $O/exp_r2f.txt 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3_final.json 2> $O/bench_c3_final.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > $O/bench_c2_final.json 2> $O/bench_c2_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_final.json 2> $O/bench_reference_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-settled > $O/ncu_bench_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vel_(divergence|update)_u" -s 8 -c 3 -o $O/r2f_pair_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-settled > $O/ncu_full_c3.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2f.txt 2>&1
cat $O/exp_r2f.txt; tail -3 $O/pytest_gpu_r2f.txt; cat $O/smoke_r2f.txt | tail -2; tail -c 300 $O/bench_c3_final.err

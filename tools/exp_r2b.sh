#!/bin/bash
# round-2 experiment B: GPU test suite on the new default build, pass-fusion / pipe-split variants, full bench lines, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > $O/pytest_gpu_r2b.txt
(SALVA_B200_REC8=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -8) > $O/pytest_gpu_r2b_rec8full.txt
(SALVA_B200_UNI_UPD=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5) > $O/pytest_gpu_r2b_alt.txt
L=salva_b200/libsalva_b200.so
echo "== variants" > $O/exp_r2b.txt
for cfg in c2 c3; do
  echo "== $cfg" >> $O/exp_r2b.txt
  timeout 1200 python tools/exp_variants.py $cfg 10 \
    default=$L \
    upd_tex=$L,SALVA_B200_UNI_UPD=1 \
    upd_alt=$L,SALVA_B200_UNI_UPD=3 \
    nofuse_akinci=$L,SALVA_B200_FUSE_AKINCI=0 \
    nofuse_xsph=$L,SALVA_B200_FUSE_XSPH=0 \
    rec8_full=$L,SALVA_B200_REC8=2 \
    fastsort=$L,BENCH_ARGS=--fast-sort \
    kernels_lib=salva_b200/libsalva_b200_kernels.so \
    >> $O/exp_r2b.txt 2>&1
done
# the driver's own invocations
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3_default.json 2> $O/bench_c3_default.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
# launch list (cold-cache, serialised: shares only) and one full capture of the pressure-iteration pair at C3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-settled > $O/ncu_bench_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vel_(divergence|update)_u" -s 8 -c 4 -o $O/r2_pair_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-settled > $O/ncu_full_c3.log 2>&1
cat $O/exp_r2b.txt; tail -3 $O/pytest_gpu_r2b.txt

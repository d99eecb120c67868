#!/bin/bash
# round-2 experiment A: pair-kernel variants (fast pair math, 256-bit gather records, g-cache, launch bounds)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/exp_r2a.txt
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_r2a.txt
L=salva_b200/variants/v_fast.so
V=salva_b200/variants
for cfg in c2 c3; do
  echo "== $cfg" >> gpurun_out/exp_r2a.txt
  timeout 1500 python tools/exp_variants.py $cfg 10 \
    base=$V/v_base.so \
    fast=$L \
    new=salva_b200/libsalva_b200.so new_nogen=$V/v_nogen.so new_rec8=salva_b200/libsalva_b200.so,SALVA_B200_REC8=1 \
    fast_rec8=$L,SALVA_B200_REC8=1 \
    fast_gcache=$L,SALVA_B200_GCACHE=1 \
    fast_gcache_rec8=$L,SALVA_B200_GCACHE=1,SALVA_B200_REC8=1 \
    m10=$V/v_m10.so m10_rec8=$V/v_m10.so,SALVA_B200_REC8=1 \
    m12=$V/v_m12.so m12_rec8=$V/v_m12.so,SALVA_B200_REC8=1 m12_gcache=$V/v_m12.so,SALVA_B200_GCACHE=1 \
    t256m4=$V/v_t256m4.so t256m4_rec8=$V/v_t256m4.so,SALVA_B200_REC8=1 \
    t256m5=$V/v_t256m5.so t256m5_rec8=$V/v_t256m5.so,SALVA_B200_REC8=1 \
    t64m18=$V/v_t64m18.so \
    fast_lsu=$L,SALVA_B200_UNI_EVAL=2,SALVA_B200_UNI_UPD=2 \
    >> gpurun_out/exp_r2a.txt 2>&1
done
cat gpurun_out/exp_r2a.txt

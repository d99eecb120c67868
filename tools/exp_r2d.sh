#!/bin/bash
# round-2 experiment D (N GPUs of one box, gpurun --gpus N): slab parity tests (NVLink peer-memory exchange, NCCL fallback,
# re-balancing) and the C4-slice bench (4M particles per GPU) with both exchange paths
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2d_gpus_$N.txt
nvidia-smi topo -m >> $O/r2d_gpus_$N.txt 2>&1
if [ "$N" = "2" ] && [ "$2" = "tests" ]; then
  (timeout 600 python -m pytest tests/test_gpu_slab.py -q 2>&1 | tail -25) > $O/pytest_gpu_2gpu.txt
fi
run() {  # name, extra env
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu_$name.json 2> $O/bench_${N}gpu_$name.err
  tail -c 400 $O/bench_${N}gpu_$name.err
}
run p2p SALVA_B200_P2P=1
if [ "$3" != "p2ponly" ]; then run nccl SALVA_B200_P2P=0; fi
python - <<PY
import json
for name in ("p2p", "nccl"):
    try:
        d = json.loads([l for l in open("$O/bench_${N}gpu_%s.json" % name) if l.startswith("{")][-1])
        print(name, "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), "slice", d.get("single_gpu_slice"), "parity", (d.get("parity") or {}).get("ok"),
              (d.get("parity") or {}).get("slab_vs_single_gpu"), "phases", {k: round(v, 3) for k, v in d["phases"].items()}, d["parallelism"])
    except Exception as e:
        print(name, "FAILED", e)
PY
cat $O/pytest_gpu_2gpu.txt 2>/dev/null | tail -5

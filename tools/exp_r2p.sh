#!/bin/bash
# 2-GPU bench only (C4 slices, peer-memory path): is the 1.48 ms grid phase of the previous run reproducible?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 95 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-parity --no-slice-ref > $O/bench_2gpu_r2p.json 2> $O/bench_2gpu_r2p.err
python - <<PY
import json
d = json.loads([l for l in open("$O/bench_2gpu_r2p.json") if l.startswith("{")][-1])
print("ms/step %.3f value %.4g grid %.3f" % (d["ms_per_step"], d["value"], d["phases"]["grid_ms"]))
print("per step", d["per_step_ms_rank0"])
PY
tail -3 $O/bench_2gpu_r2p.err

#!/bin/bash
# 2-GPU validation of the sharded step (both exchange forms) + weak-scaling bench at N=2
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q) > gpurun_out/tm.log 2>&1; tail -15 gpurun_out/tm.log
for comm in nccl p2p; do
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --comm $comm) > gpurun_out/bm_$comm.log 2> gpurun_out/bm_$comm.err
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/bm_$comm.log").read().strip().splitlines()[-1])
    print("$comm", d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["parity_check"])
except Exception as e:
    print("$comm failed", e); print(open("gpurun_out/bm_$comm.err").read()[-3000:])
EOF
done

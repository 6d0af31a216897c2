#!/bin/bash
# 8-GPU validation: weak scaling with the in-kernel peer exchange (+ config 5 row-sharded), then the NCCL step for comparison
mkdir -p gpurun_out
N=${1:-8}
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 20 --warmup 5 --comm p2p --c5 on) > gpurun_out/b8_p2p.log 2> gpurun_out/b8_p2p.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus $N --steps 200 --warmup 20 --comm p2p --c5 off) > gpurun_out/b8_p2p200.log 2> gpurun_out/b8_p2p200.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus $N --steps 20 --warmup 5 --comm nccl --c5 off) > gpurun_out/b8_nccl.log 2> gpurun_out/b8_nccl.err
python - <<EOF
import json
for name in ("b8_p2p", "b8_p2p200", "b8_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/{name}.log").read().strip().splitlines()[-1])
        print(name, "value", d["value"]/1e9, "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, "parity", d["parity_check"]["ok"], d["parity_check"]["max_rel_loss"], d["parity_check"]["max_abs_table"])
        if "configs" in d: print("   c5", json.dumps(d["configs"]["c5"])[:600])
    except Exception as e:
        print(name, "failed", e); print(open(f"gpurun_out/{name}.err").read()[-2500:])
EOF

#!/bin/bash
# N-GPU validation (default 8).  CHARGED N x box time: every command carries its own short timeout, and the cheap probes come
# first, so that a stage that blocks costs minutes, not the round's budget (round 2 lost ~145 GPU-minutes to one 900 s timeout).
#   1. NCCL step, 20 steps, no config 5          (the default at N > 2)
#   2. peer-exchange kernel, 20 steps            (--comm p2p; bench.py's own watchdog prints what it has after 150 s)
#   3. peer-exchange kernel, 200 steps           (only if 2. produced a value)
#   4. NCCL step with config 5                   (Netflix shape, F = 128, row-sharded)
mkdir -p gpurun_out
N=${1:-8}
run() {  # name timeout args...
  local name=$1 to=$2; shift 2
  (timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
     bench.py --gpus $N "$@") > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "== $name rc=$?"; tail -c 1500 gpurun_out/$name.log; tail -c 600 gpurun_out/$name.err
}
run b${N}_nccl 300 --steps 20 --warmup 5 --comm nccl --c5 off --watchdog 200
run b${N}_p2p 240 --steps 20 --warmup 5 --comm p2p --c5 off --watchdog 150
if grep -q '"value": [0-9]' gpurun_out/b${N}_p2p.log && ! grep -q incomplete gpurun_out/b${N}_p2p.log; then
  run b${N}_p2p200 240 --steps 200 --warmup 20 --comm p2p --c5 off --watchdog 150
fi
run b${N}_nccl_c5 420 --steps 20 --warmup 5 --comm nccl --c5 on --watchdog 330

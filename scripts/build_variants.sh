#!/bin/bash
# dev tool: build A/B variants of the step kernel (launch bounds / unroll) into gpurun_out-free scratch libs
set -e
cd "$(dirname "$0")/.."
python daisyrec_b200/_build.py >/dev/null
mkdir -p daisyrec_b200/lib/variants
for cfg in "$@"; do
  mb=${cfg%%_*}; un=${cfg##*_}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DDRB_MINB=$mb -DDRB_UNR=$un \
       -c daisyrec_b200/csrc/mf_bpr.cu -o daisyrec_b200/lib/variants/mf_bpr_${cfg}.o &
done
wait
for cfg in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o daisyrec_b200/lib/variants/lib_${cfg}.so \
       daisyrec_b200/lib/variants/mf_bpr_${cfg}.o daisyrec_b200/lib/capi.o daisyrec_b200/lib/sampler.o daisyrec_b200/lib/rank.o daisyrec_b200/lib/shard.o daisyrec_b200/lib/lightgcn.o daisyrec_b200/lib/neumf.o daisyrec_b200/lib/comm.o daisyrec_b200/lib/metrics.o -ldl
  rm daisyrec_b200/lib/variants/mf_bpr_${cfg}.o
done
ls daisyrec_b200/lib/variants

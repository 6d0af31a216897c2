#!/bin/bash
# developer probe: MT19937 stream kernel thread-count variants at n = 80 M (run on the GPU box)
for t in 32 64 128 256; do
  DRB_MT_THREADS=$t python - <<EOF
import torch, sys
sys.path.insert(0, ".")
from daisyrec_b200 import ops
import bench
n = 80_000_000
ms = bench.timed_ms(lambda: ops.mt19937_stream(2022, n, "cuda"), 1, 3)
print("DRB_MT_THREADS=$t", round(ms, 2), "ms")
EOF
done

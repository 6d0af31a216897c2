#!/bin/bash
# Round-2 single-GPU validation: step-kernel A/B, the whole GPU suite, the driver's bench line, the reference arm.
mkdir -p gpurun_out
{
for v in "DRB_NO_LEAN=1" "DRB_LEAN_NCH=1" "DRB_LEAN_NCH=2" "DRB_LEAN_NCH=4" "DRB_LEAN_NCH=4 DRB_TILE_CAP=512"; do
  echo "== $v"; env $v timeout 300 python scripts/ab_step.py c2 c5
done
echo "== adam (default lean)"; timeout 300 python scripts/ab_step.py c2 adam
echo "== adam (general)"; DRB_NO_LEAN=1 timeout 300 python scripts/ab_step.py c2 adam
} > gpurun_out/ab_step.log 2>&1
cat gpurun_out/ab_step.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py 2>&1 | tail -15 | tee gpurun_out/r02_gpu_suite.log
if [ "$1" == "bench" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
  tail -c 6000 gpurun_out/r02_bench_n1.json
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
  cat gpurun_out/r02_bench_ref.json
fi

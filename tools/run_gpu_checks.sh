#!/bin/bash
# Run every GPU test file in its own process (a device trap is sticky per process) and keep the logs.
# usage: tools/run_gpu_checks.sh [files...]
mkdir -p gpurun_out
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_gpu_gemm.py tests/test_gpu_gemm_pair.py tests/test_gpu_attention_tc.py tests/test_gpu_ops.py tests/test_gpu_video.py tests/test_gpu_e2e.py); fi
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "${files[@]}"; do
  b=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -m gpu -q -s --no-header --tb=short -p no:cacheprovider > "gpurun_out/$b.log" 2>&1
  r=$?
  echo "$f -> exit $r"
  tail -n 40 "gpurun_out/$b.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc

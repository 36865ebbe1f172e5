#!/bin/bash
# Run the GPU test-suite on the box file by file (own timeout each) and keep the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${@:-tests/test_checksum_gpu.py tests/test_inflate_gpu.py tests/test_deflate_gpu.py}; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -x -q > gpurun_out/$b.log 2>&1
  echo "exit $?" >> gpurun_out/$b.log
  echo "=== $b"; tail -25 gpurun_out/$b.log
done

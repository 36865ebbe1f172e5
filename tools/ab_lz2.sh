#!/bin/bash
# throughput-only A/B of lz77_kernel builds (run on the GPU box): each argument is "label|nvcc flags|env settings"; with PROF=1 the
# role-cycle dump of -DZS_LZ_PROF (tools/lzprof.py) is taken as well.  Output: gpurun_out/ab_lz2.txt
P=zlib-streams-ts_b200
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
for spec in "$@"; do
  IFS='|' read -r label flags envs <<< "$spec"
  ZS_NVCC_EXTRA="$flags" python $P/build.py --force > /dev/null || { echo "build failed [$flags]"; continue; }
  echo "=== $label [flags: $flags] [env: $envs]"
  env $envs timeout 300 python tools/defprof.py 2>&1 | cut -c1-170
  if [ -n "$PROF" ]; then
    ZS_NVCC_EXTRA="$flags -DZS_LZ_PROF" python $P/build.py --force > /dev/null && env $envs timeout 120 python tools/lzprof.py 2>&1 | grep -E "prof\]|level" | cut -c1-260
  fi
done
python $P/build.py --force > /dev/null
} > gpurun_out/ab_lz2.txt 2>&1
cat gpurun_out/ab_lz2.txt

#!/bin/bash
# A/B driver for lz77_kernel policies (run on the GPU box).  Each argument is "label|nvcc flags|env settings":
# the library is rebuilt with the flags, then the size table (tools/ratiocheck.py) and the throughput table
# (tools/defprof.py) run with the environment settings.  Output: gpurun_out/ab_lz.txt
#   e.g. tools/ab_lz.sh "base||" "scanend|-DZS_LZ_SCANEND|" "a12||ZS_LZ_STOP_ACTIVE=12"
LEVELS=${AB_LEVELS:-"6 9"}
P=zlib-streams-ts_b200
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
last="?"
for spec in "$@"; do
  IFS='|' read -r label flags envs <<< "$spec"
  if [ "$flags" != "$last" ]; then
    ZS_NVCC_EXTRA="$flags" python $P/build.py --force > /dev/null || { echo "build failed [$flags]"; continue; }
    last="$flags"
  fi
  echo "=== $label [flags: $flags] [env: $envs]"
  env $envs timeout 600 python tools/ratiocheck.py $LEVELS 2>&1 | sed 's/gpu.zlib same plan (gpu.zlib one shot)://' | cut -c1-150
  env $envs timeout 600 python tools/defprof.py 2>&1 | cut -c1-150
done
python $P/build.py --force > /dev/null
} > gpurun_out/ab_lz.txt 2>&1
tail -n 100 gpurun_out/ab_lz.txt

#!/bin/bash
# A/B build variants: ratio vs zlib (same plan) over data kinds at the given levels + L6 speed on text / runs
P=zlib-streams-ts_b200
LV="$1"; shift
for v in "$@"; do
  ZS_NVCC_EXTRA="$v" python $P/build.py --force > /dev/null || { echo "build failed [$v]"; continue; }
  echo "=== [$v]"; python tools/ratiocheck.py $LV 2>&1 | sed 's/gpu.zlib same plan (gpu.zlib one shot)://' | cut -c1-110
  python tools/tileprof.py 6 2>&1 | grep "^text\|^runs"
done
python $P/build.py --force > /dev/null

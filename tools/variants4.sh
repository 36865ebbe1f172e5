#!/bin/bash
# A/B build variants for the greedy levels: ratio vs zlib (same plan) at the given levels + bench headline
P=zlib-streams-ts_b200
LV="$1"; shift
for v in "$@"; do
  ZS_NVCC_EXTRA="$v" python $P/build.py --force > /dev/null || { echo "build failed [$v]"; continue; }
  echo "=== [$v]"; python tools/ratiocheck.py $LV 2>&1 | sed 's/gpu.zlib same plan (gpu.zlib one shot)://' | cut -c1-130
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['compressed_ratio'])"
done
python $P/build.py --force > /dev/null

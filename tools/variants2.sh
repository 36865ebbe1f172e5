#!/bin/bash
# A/B build variants: bench extra (L6 ratio/speed) + per-kind throughput
P=zlib-streams-ts_b200
for v in "$@"; do
  ZS_NVCC_EXTRA="$v" python $P/build.py --force > /dev/null || { echo "build failed [$v]"; continue; }
  echo "=== [$v]"; python bench.py --steps 2 --warmup 3 --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['compressed_ratio'], d['extra'])"
  python tools/tileprof.py 6 9 2>&1 | grep -v "^ramp\|^zeros\|^random"
done

#!/bin/bash
# A/B build variants of the LZ77 kernel on the GPU box: role cycles per variant (tools/lzprof.py)
P=zlib-streams-ts_b200
for v in "$@"; do
  ZS_NVCC_EXTRA="-DZS_LZ_PROF $v" python $P/build.py --force > /dev/null || { echo "build failed [$v]"; continue; }
  echo "=== [$v]"; python tools/lzprof.py 2>&1 | grep "lz77 prof" | cut -c1-120
done

#!/bin/bash
# A/B for the lazy levels' straggler stop (search_position_sync, -DZS_LZ_STRAGGLER): size against zlib over
# the eight data kinds at levels 6 and 9, and throughput of the deflate table (tools/defprof.py), for the
# default build and for a few (ZS_STOP_AFTER, ZS_STOP_ACTIVE) pairs.  The CPU model's predictions for the
# same pairs: python tools/lzmodel.py 6 9 -- stop_active=K stop_after=M
P=zlib-streams-ts_b200
for v in "" "-DZS_LZ_STRAGGLER" "-DZS_LZ_STRAGGLER -DZS_STOP_ACTIVE=2" "-DZS_LZ_STRAGGLER -DZS_STOP_ACTIVE=8" "-DZS_LZ_STRAGGLER -DZS_STOP_AFTER=16"; do
  ZS_NVCC_EXTRA="$v" python $P/build.py --force > /dev/null || { echo "build failed [$v]"; continue; }
  echo "=== [$v]"
  timeout 600 python tools/ratiocheck.py 6 9 2>&1 | sed 's/gpu.zlib same plan (gpu.zlib one shot)://' | cut -c1-130
  timeout 600 python tools/defprof.py 2>&1 | grep -E "L6|L9" | cut -c1-110
  timeout 300 python -m pytest tests/test_deflate_gpu.py -x -q -m gpu 2>&1 | tail -1
done
python $P/build.py --force > /dev/null

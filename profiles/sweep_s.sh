#!/bin/bash
# streams-per-warp sweep on config 2 (10k reads)
for S in 4 8 16 32; do
  GA_STREAMS_PER_WARP=$S python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('S=$S', 'kernel_ms', round(d['ms_per_step'],2), 'gcups', round(d['gcups'],1))"
  GA_DEBUG_FLAGS=1 GA_STREAMS_PER_WARP=$S python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   forward only', round(d['ms_per_step'],2))"
done

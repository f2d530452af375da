#!/bin/bash
# batch-size study: R copies of config 2's 10k-read set per step, and streams-per-warp overrides
run() {
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > /tmp/sweep_out.txt 2> /tmp/sweep_err.txt
  if ! tail -1 /tmp/sweep_out.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$LABEL', 'kernel_ms', round(d['ms_per_step'],2), 'gcups', round(d['gcups'],1), 'Gbp/s', round(d['value']/1e9,2), 'e2e_ms', round(d['e2e']['ms_per_step'],1), 'int_frac', round(d['roofline']['int_alu']['frac'],4))" 2>/dev/null; then
    echo "$LABEL FAILED: $(grep -v 'nodes$\|bp$\|edges$\|in-degree' /tmp/sweep_err.txt | tail -1 | cut -c1-200)"
  fi
}
LABEL="R1_auto(S4)" run
LABEL="R4_auto" run --replicate 4
LABEL="R4_S16" GA_STREAMS_PER_WARP=16 run --replicate 4
LABEL="R4_S32" GA_STREAMS_PER_WARP=32 run --replicate 4
LABEL="R8_auto" run --replicate 8
LABEL="R8_S16" GA_STREAMS_PER_WARP=16 run --replicate 8

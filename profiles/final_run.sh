#!/bin/bash
# what the round-end evidence in profiles/ comes from (1x B200): tests, smoke, both bench arms, ncu launch list
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_ref.json")); o = json.load(open("gpurun_out/bench_own.json"))
print("reference Mbp/s", r["value"] / 1e6, "| kernel ms", o["ms_per_step"], "Gbp/s", o["value"] / 1e9, "GCUPS", o["gcups"],
      "| e2e ms", o["e2e"]["ms_per_step"], "Gbp/s", o["e2e"]["value"] / 1e9, "| hbm frac", o["roofline"]["frac"], "int frac", o["roofline"]["int_alu"]["frac"])
PY
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
grep -c ga_align gpurun_out/launches.csv

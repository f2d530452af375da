cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --stock-cpu > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r02_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_r02_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ga_fast_kernel -s 1 -c 1 -o gpurun_out/r02c_fast -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
ncu --set full --clock-control none --import-source on -k regex:ga_trace_kernel -s 1 -c 1 -o gpurun_out/r02c_trace -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
run() { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.3g bp/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'gcups %.0f' % d['gcups'], 'e2e %.3g bp/s %.2f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'split', d['kernel_split_ms'], 'failed', d['failed_reads'], 'prep %.0fs' % d['prep_s'])
    print('   parity', d.get('parity')); print('   cpu', d.get('cpu_baseline'))
    for m in d.get('sweep', []): print('   band', m['band'], 'value %.3g' % m['value'], 'gcups %.0f' % m['gcups'], 'e2e %.3g' % m['e2e']['value'], m['kernel_split_ms'], 'rerun', m['streams_rerun_with_general_layout'])
except Exception as e: print('$name: no line', e)
PY
}
run bench_r02_config5 --config 5 --steps 3 --warmup 3
run bench_r02_config3 --config 3 --steps 3 --warmup 3

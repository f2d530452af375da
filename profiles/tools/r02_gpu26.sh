cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/tools/many_seeds_probe.py 1000 40 2>&1 | tail -1
run() { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.3g bp/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'gcups %.0f' % d['gcups'], 'e2e %.3g bp/s %.2f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'split', d['kernel_split_ms'], 'failed', d['failed_reads'], 'prep %.0fs' % d['prep_s'], 'd2h', d['e2e']['d2h_bytes_per_step'], 'dev', d['e2e']['device_ms_per_step'])
    print('   parity', d.get('parity')); print('   cpu', d.get('cpu_baseline'))
except Exception as e: print('$name: no line', e)
PY
}
run bench_r02_config3 --config 3 --steps 3 --warmup 3

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2
run() { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -3 gpurun_out/$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.3g bp/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'gcups %.0f' % d['gcups'], 'e2e %.3g bp/s %.2f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'split', d['kernel_split_ms'], 'roof', d['roofline']['frac'], 'failed', d['failed_reads'], 'prep %.0fs' % d['prep_s'])
    print('   parity', d.get('parity'))
    print('   cpu', d.get('cpu_baseline'), d.get('cpu_baseline_stock'))
    for m in d.get('sweep', []): print('   band', m['band'], 'value %.3g' % m['value'], 'gcups %.0f' % m['gcups'], 'e2e %.3g' % m['e2e']['value'], m['kernel_split_ms'], 'failed', m['failed_reads'], 'err', m['stream_errors'], 'streams', m['streams'], 'rerun', m['streams_rerun_with_general_layout'])
except Exception as e: print('$name: no line', e)
PY
}
run bench_r02_config3 --config 3 --steps 3 --warmup 3
run bench_r02_config5 --config 5 --steps 3 --warmup 3

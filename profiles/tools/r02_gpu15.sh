cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 1500 gpurun_out/$name.err | tail -5; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.3g bp/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'gcups %.0f' % d['gcups'], 'e2e %.3g bp/s %.2f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'split', d['kernel_split_ms'], 'roof', d['roofline']['frac'], 'failed', d['failed_reads'], 'prep %.0fs' % d['prep_s'])
    print('   parity', d.get('parity'))
    print('   cpu', d.get('cpu_baseline'), d.get('cpu_baseline_stock'))
    for m in d.get('sweep', []): print('   band', m['band'], 'value %.3g' % m['value'], 'gcups %.0f' % m['gcups'], 'e2e %.3g' % m['e2e']['value'], m['kernel_split_ms'], 'failed', m['failed_reads'], 'err', m['stream_errors'])
except Exception as e: print('$name: no line', e)
PY
}
run t_c2 --steps 3 --warmup 3 --stock-cpu --cpu-sample 400
run t_c3 --config 3 --scale 0.02 --steps 2 --warmup 3 --parity-sample 16
run t_c5 --config 5 --scale 0.02 --steps 2 --warmup 3 --parity-sample 4
run t_c4 --config 4 --scale 0.002 --reads 1000 --steps 2 --warmup 3 --parity-sample 8
run t_ref3 --impl reference --config 3 --scale 0.02 --steps 1 --warmup 0 --cpu-sample 16

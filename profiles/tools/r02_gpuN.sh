# usage: r02_gpuN.sh N "<config-4 args>"   (N = GPUs of the box)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
nproc; nvidia-smi -L | wc -l
run() { name=$1; shift; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('$name', 'n_gpus', d['n_gpus'], 'value %.3g bp/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'gcups %.0f' % d['gcups'], 'e2e %.3g bp/s %.2f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'split', d['kernel_split_ms'], 'failed', d['failed_reads'], 'prep %.0fs' % d['prep_s'])
    print('   parity', d.get('parity'))
except Exception as e:
    print('$name: no line', e); import subprocess; print(subprocess.run(['tail','-5','gpurun_out/$name.err'],capture_output=True,text=True).stdout)
PY
}
run bench_r02_${N}gpu --steps 8 --warmup 3
if [ -n "$2" ]; then run bench_r02_config4_${N}gpu --config 4 $2; fi

# usage: r02_gpuN2.sh N   - config 2 at N GPUs, with the host stage timers of rank 0
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
nproc
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; grep "ga timing" gpurun_out/$name.err | grep -v "free" | tail -24 | sort | uniq -c | sort -rn | head -14; python - <<PY
import json
d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
print('$name', 'n_gpus', d['n_gpus'], 'kernel ms/step %.2f' % d['ms_per_step'], 'e2e %.3g bp/s %.2f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'arrivals', d['e2e']['batch_arrival_ms'])
PY
}
run q_${N}_block GA_TIMING=1
run q_${N}_spin GA_TIMING=1 GA_SPIN_WAIT=1
run q_${N}_pageable GA_TIMING=1 GA_BENCH_PAGEABLE=1

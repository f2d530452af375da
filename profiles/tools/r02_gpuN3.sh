cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
GA_TIMELINE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/tl_$N.json 2> gpurun_out/tl_$N.err; echo rc=$?
python - <<PY
import json,re,collections
d=json.loads(open('gpurun_out/tl_$N.json').read().strip().splitlines()[-1])
print('e2e %.2f ms' % d['e2e']['ms_per_step'], d['e2e']['batch_arrival_ms'])
PY
grep "ga timeline" gpurun_out/tl_$N.err | wc -l

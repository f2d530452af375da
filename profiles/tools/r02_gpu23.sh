cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for TH in 16 4; do
GA_HOST_THREADS=$TH GA_TIMING=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p_$TH.json 2> gpurun_out/p_$TH.err
grep "ga timing" gpurun_out/p_$TH.err | grep -v "free\|d2h:" | tail -7
python - <<PY
import json
d=json.loads(open('gpurun_out/p_$TH.json').read().strip().splitlines()[-1])
print('threads $TH: kernel ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'single', round(d['e2e']['single_call_ms'],2), d['kernel_split_ms'], d['e2e']['batch_arrival_ms'])
PY
done

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -3
GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/n.json 2> gpurun_out/n.err
echo "c2: $(grep 'ga kernels' gpurun_out/n.err | sed -n 6p)"
python - <<PY
import json
d=json.loads(open('gpurun_out/n.json').read().strip().splitlines()[-1])
print('kernel ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'single', round(d['e2e']['single_call_ms'],2), d['kernel_split_ms'], d['e2e']['batch_arrival_ms'])
PY

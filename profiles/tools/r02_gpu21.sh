cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/y.json 2> gpurun_out/y.err
echo "c2: $(grep 'ga kernels' gpurun_out/y.err | sed -n 6p)"
GA_NO_L2_WINDOW=1 GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/y2.json 2> gpurun_out/y2.err
echo "c2 no L2 window: $(grep 'ga kernels' gpurun_out/y2.err | sed -n 6p)"
python - <<PY
import json
d=json.loads(open('gpurun_out/y.json').read().strip().splitlines()[-1])
print('   value %.3g ms %.2f e2e %.3g (%.1f ms/step) single %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_call_ms']), d['kernel_split_ms'], 'roof', d['roofline']['frac'])
PY
GA_KERNEL_TIMING=1 timeout 600 python bench.py --config 3 --scale 0.1 --reads 10000 --parity-sample 0 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/y3.json 2> gpurun_out/y3.err
echo "c3: $(grep 'ga kernels' gpurun_out/y3.err | sed -n 5p)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3

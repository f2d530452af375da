cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/w.json 2> gpurun_out/w.err
echo "c2: $(grep 'ga kernels' gpurun_out/w.err | sed -n 6p)"
python - <<PY
import json
d=json.loads(open('gpurun_out/w.json').read().strip().splitlines()[-1])
print('   value %.3g ms %.2f e2e %.3g (%.1f ms/step) single %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_call_ms']), d['kernel_split_ms'], 'roof', d['roofline']['frac'])
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3

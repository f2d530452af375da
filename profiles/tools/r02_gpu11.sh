cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
GA_TIMING=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b11.json 2> gpurun_out/b11.err
grep "ga timing" gpurun_out/b11.err | tail -12
python - <<PY
import json
d=json.loads(open('gpurun_out/b11.json').read().strip().splitlines()[-1])
print('kernel ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'single', round(d['e2e']['single_call_ms'],2), 'h2d', d['e2e']['h2d_bytes_per_step'], 'd2h', d['e2e']['d2h_bytes_per_step'], 'failed', d['failed_reads'])
PY
for D in 2 3; do
GA_PIPELINE_DEPTH=$D GA_HOST_THREADS=4 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b11_D$D.json 2> gpurun_out/b11_D$D.err
python - <<PY
import json
d=json.loads(open('gpurun_out/b11_D$D.json').read().strip().splitlines()[-1])
print('4 host threads, depth $D: kernel ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'single', round(d['e2e']['single_call_ms'],2))
PY
done

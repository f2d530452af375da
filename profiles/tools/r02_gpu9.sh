cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for D in 2 3; do for FR in 0 1; do
if [ $FR = 1 ]; then export GA_PIPELINE_FREE_RUN=1; else unset GA_PIPELINE_FREE_RUN; fi
GA_PIPELINE_DEPTH=$D timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b9_D${D}_F$FR.json 2> gpurun_out/b9_D${D}_F$FR.err
python - <<PY
import json
d=json.loads(open('gpurun_out/b9_D${D}_F$FR.json').read().strip().splitlines()[-1])
print('depth $D free $FR: kernel ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'single', round(d['e2e']['single_call_ms'],2))
PY
done; done

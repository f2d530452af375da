cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for TH in 16 4; do
GA_HOST_THREADS=$TH GA_TIMING=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/m_$TH.json 2> gpurun_out/m_$TH.err
grep "ga timing" gpurun_out/m_$TH.err | grep -v "free\|d2h:" | tail -7
python - <<PY
import json
d=json.loads(open('gpurun_out/m_$TH.json').read().strip().splitlines()[-1])
print('threads $TH: kernel ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'single', round(d['e2e']['single_call_ms'],2), d['kernel_split_ms'], 'd2h', d['e2e']['d2h_bytes_per_step'], d['e2e']['batch_arrival_ms'])
PY
done
timeout 600 python profiles/tools/full_parity.py > gpurun_out/r02_config2_full_parity.txt 2> gpurun_out/full_parity.err; echo "full parity rc=$?"; tail -4 gpurun_out/r02_config2_full_parity.txt

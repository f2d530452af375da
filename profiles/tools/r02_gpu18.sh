cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for T in 4 5 6 7 8; do
GA_TRACE_T=$T GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/u_$T.json 2> gpurun_out/u_$T.err
echo "c2 T=$T: $(grep 'ga kernels' gpurun_out/u_$T.err | sed -n 5p)"
done
for S in 16 32; do for T in 6 8 16 32; do
GA_STREAMS_PER_WARP=$S GA_TRACE_T=$T GA_KERNEL_TIMING=1 timeout 600 python bench.py --config 3 --scale 0.1 --reads 10000 --parity-sample 0 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/v_${S}_$T.json 2> gpurun_out/v_${S}_$T.err
echo "c3 S=$S T=$T: $(grep 'ga kernels' gpurun_out/v_${S}_$T.err | sed -n 5p)"
python - <<PY
import json
d=json.loads(open('gpurun_out/v_${S}_$T.json').read().strip().splitlines()[-1])
print('   value %.3g e2e %.3g (%.1f ms/step) device/step' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']), d['e2e']['device_ms_per_step'], 'streams', d.get('streams'))
PY
done; done

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
one() { name=$1; shift; env "$@" GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name: $(grep 'ga kernels' gpurun_out/$name.err | sed -n 6p)"; }
one z_main X=1
python - <<PY
import json
d=json.loads(open('gpurun_out/z_main.json').read().strip().splitlines()[-1])
print('   value %.3g ms %.2f e2e %.3g (%.1f ms/step) single %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_call_ms']), d['kernel_split_ms'])
print('   arrivals', d['e2e']['batch_arrival_ms'])
PY
one z_a16_T6 GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt16.so GA_TRACE_T=6
one z_a16_T5 GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt16.so GA_TRACE_T=5
one z_a16_T4 GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt16.so GA_TRACE_T=4
one z_a14_T5 GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt14.so GA_TRACE_T=5
one z_a14_T6 GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt14.so GA_TRACE_T=6

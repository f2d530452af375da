cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for S in 16 17 18 20 24; do
GA_STREAMS_PER_WARP=$S GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s_$S.json 2> gpurun_out/s_$S.err
echo "S=$S: $(grep 'ga kernels' gpurun_out/s_$S.err | sed -n 5p)"
done
for T in 6 10 12; do
GA_STREAMS_PER_WARP=17 GA_TRACE_T=$T GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/t_$T.json 2> gpurun_out/t_$T.err
echo "S=17 T=$T: $(grep 'ga kernels' gpurun_out/t_$T.err | sed -n 5p)"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3

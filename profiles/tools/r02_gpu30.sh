cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v2.json 2> gpurun_out/v2.err
echo "c2: $(grep 'ga kernels' gpurun_out/v2.err | sed -n 5,7p)"

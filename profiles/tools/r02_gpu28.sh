cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/u2.json 2> gpurun_out/u2.err
echo "c2: $(grep 'ga kernels' gpurun_out/u2.err | sed -n 6p)"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -3

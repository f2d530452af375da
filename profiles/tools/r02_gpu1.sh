set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for S in 4 8 16; do for G in 8 32; do
GA_KERNEL_TIMING=1 GA_STREAMS_PER_WARP=$S GA_TRACE_GROUP=$G timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_S${S}_G${G}.json 2> gpurun_out/bench_S${S}_G${G}.err; grep "ga kernels" gpurun_out/bench_S${S}_G${G}.err | head -3
done; done
GA_KERNEL_TIMING=1 GA_TRACE_GROUP=4 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_G4.json 2> gpurun_out/bench_G4.err; grep "ga kernels" gpurun_out/bench_G4.err | head -3

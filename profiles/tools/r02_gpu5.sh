cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for S in 8 16; do
GA_KERNEL_TIMING=1 GA_STREAMS_PER_WARP=$S timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b5_S$S.json 2> gpurun_out/b5_S$S.err
grep -E "ga kernels" gpurun_out/b5_S$S.err | tail -2
done
bash profiles/tools/r02_ncu1.sh

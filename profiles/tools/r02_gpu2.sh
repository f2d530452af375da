cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for S in 4 16; do
GA_LIB=$PWD/graphaligner_b200/libga_phase.so GA_KERNEL_TIMING=1 GA_STREAMS_PER_WARP=$S timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/phase_S$S.json 2> gpurun_out/phase_S$S.err
grep -E "ga kernels|ga phases" gpurun_out/phase_S$S.err | tail -20
done

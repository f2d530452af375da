cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for T in 4 8 16 32; do for P in 1 2; do
GA_KERNEL_TIMING=1 GA_TRACE_T=$T GA_TRACE_P=$P timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b6_T${T}_P$P.json 2> gpurun_out/b6_T${T}_P$P.err
echo "T $T P $P: $(grep -E 'ga kernels' gpurun_out/b6_T${T}_P$P.err | tail -1)"
done; done

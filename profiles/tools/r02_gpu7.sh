cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for T in 4 8 16; do for P in 1 2; do
GA_KERNEL_TIMING=1 GA_TRACE_T=$T GA_TRACE_P=$P timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b7_T${T}_P$P.json 2> gpurun_out/b7_T${T}_P$P.err
echo "T $T P $P: $(grep -E 'ga kernels' gpurun_out/b7_T${T}_P$P.err | tail -1)"
done; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
GA_TRACE_T=8 GA_TRACE_P=2 ncu --set full --clock-control none --import-source on -k regex:ga_trace_kernel -s 1 -c 1 -o gpurun_out/r02_trace5_T8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for T in 8 32; do
GA_TRACE_T=$T GA_TRACE_P=1 ncu --set full --clock-control none --import-source on -k regex:ga_trace_kernel -s 1 -c 1 -o gpurun_out/r02_trace4_T$T -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
done
ls -la gpurun_out/*.ncu-rep

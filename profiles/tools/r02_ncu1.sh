cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ga_trace_kernel -s 1 -c 1 -o gpurun_out/r02_trace3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
ls -la gpurun_out/*.ncu-rep

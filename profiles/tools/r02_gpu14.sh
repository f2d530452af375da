cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err || exit 1
cat gpurun_out/r02b_bench.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:ga_fast_kernel -s 1 -c 1 -o gpurun_out/r02b_fast -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
ncu --set full --clock-control none --import-source on -k regex:ga_trace_kernel -s 1 -c 1 -o gpurun_out/r02b_trace -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
ls -la gpurun_out/r02b*

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r02_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200

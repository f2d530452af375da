cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/w2.json 2> gpurun_out/w2.err
echo "c2: $(grep 'ga kernels' gpurun_out/w2.err | sed -n 6p)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/w2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
grep "ga_peq_kernel\|ga_validate_kernel" gpurun_out/w2_launches.csv | tail -4 | cut -d, -f5,15-

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
GA_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/b10.json 2> gpurun_out/b10.err
grep "ga timing" gpurun_out/b10.err | tail -40

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
one() { name=$1; shift; env "$@" GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name: $(grep 'ga kernels' gpurun_out/$name.err | sed -n 6p)"; }
one s2 X=1
one s3 GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt_k3.so
one s4 GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt_k4.so

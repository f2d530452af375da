cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
one() { name=$1; shift; env "$@" GA_KERNEL_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name: $(grep 'ga kernels' gpurun_out/$name.err | sed -n 6p)"; }
one k_main X=1
one k_pf GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt_pf.so
one k_lb GA_LIB=$GRAFT_REPO_ROOT/graphaligner_b200/libga_alt_lb.so

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5

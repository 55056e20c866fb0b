cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -m gpu -x -q 2>&1 | tail -4

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair" > gpurun_out/r2_pytest_pair2.log 2>&1; tail -8 gpurun_out/r2_pytest_pair2.log

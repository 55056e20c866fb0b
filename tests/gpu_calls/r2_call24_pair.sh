cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B200_FA_PAIR=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fa_ or attention" > gpurun_out/r2_pytest_pair.log 2>&1; tail -12 gpurun_out/r2_pytest_pair.log
B200_FA_PAIR=1 timeout 120 python tests/attn_quick.py std 2>&1 | tail -4
timeout 120 python tests/attn_quick.py std 2>&1 | tail -4
VARIANTS="old pair" timeout 900 bash tests/fa_ab3.sh 2>&1 | tail -12

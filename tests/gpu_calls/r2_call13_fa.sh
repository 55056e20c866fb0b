set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fa_ or attention or paged_short" > gpurun_out/r2_pytest_fa.log 2>&1; tail -5 gpurun_out/r2_pytest_fa.log
timeout 120 python tests/attn_quick.py std 2>&1 | tail -8
VARIANTS="old new" timeout 900 bash tests/fa_ab3.sh 2>&1 | tail -14
FA_AB_D=64 VARIANTS="old new" timeout 900 bash tests/fa_ab3.sh 2>&1 | tail -14

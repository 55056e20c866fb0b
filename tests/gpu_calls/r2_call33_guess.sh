cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fa_ or attention or attn" 2>&1 | tail -3
VARIANTS="old guess" timeout 900 bash tests/fa_ab3.sh 2>&1 | tail -11
FA_AB_D=64 VARIANTS="old guess" timeout 900 bash tests/fa_ab3.sh 2>&1 | tail -3
timeout 120 python tests/attn_quick.py std 2>&1 | tail -3

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
VARIANTS="old pair pairp0 pairp8" timeout 900 bash tests/fa_ab3.sh 2>&1 | tail -13

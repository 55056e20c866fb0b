cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
VARIANTS="c2 c7" timeout 1200 bash tests/fa_ab3.sh 2>&1 | tail -16

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
VARIANTS="old v1 v2 v12" timeout 900 bash tests/fa_ab3.sh 2>&1 | tail -13

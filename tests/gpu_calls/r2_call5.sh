set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo skip probe

( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest2.log 2>&1
tail -15 gpurun_out/r2_pytest2.log
cat gpurun_out/fullsize_parity.json

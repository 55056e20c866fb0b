set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/fused_mlp_probe.py > gpurun_out/r2_fused_probe.log 2>&1; echo rc=$?
tail -20 gpurun_out/r2_fused_probe.log

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python tests/fused_mlp_probe.py > gpurun_out/r2_fused_probe.log 2>&1; echo rc=$?
tail -8 gpurun_out/r2_fused_probe.log
timeout 120 python tests/ncu_target.py mlp 2 > gpurun_out/ncu_plain_mlp.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_mlp -c 1 -o gpurun_out/prof_fusedmlp_c3_r2 python tests/ncu_target.py mlp 2 > gpurun_out/ncu_fusedmlp_c3.log 2>&1
timeout 120 python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_plain_mlp2.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_mlp -c 1 -o gpurun_out/prof_fusedmlp_c2_r2 python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_fusedmlp_c2.log 2>&1
ls -la gpurun_out | tail -5

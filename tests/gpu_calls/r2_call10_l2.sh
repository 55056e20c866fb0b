set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/gemm_l2_probe.py > gpurun_out/r2_gemm_l2_time.log 2>&1; echo rc=$?
cat gpurun_out/r2_gemm_l2_time.log | tail -3
NCU=1 timeout 200 python tests/gemm_l2_probe.py > gpurun_out/ncu_plain_l2.log 2>&1 && NCU=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control none -k regex:gemm_act_pair --csv --log-file gpurun_out/r2_gemm_l2_ncu.csv python tests/gemm_l2_probe.py > gpurun_out/ncu_l2.log 2>&1
tail -2 gpurun_out/ncu_l2.log
timeout 200 python tests/ln_probe.py > gpurun_out/r2_ln_probe.log 2>&1; cat gpurun_out/r2_ln_probe.log | head -24

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/ncu_cudnn_target.py > gpurun_out/ncu_plain_cudnn.log 2>&1 && timeout 600 ncu --set full --clock-control none -s 1 -c 1 -k regex:sdpa -o gpurun_out/r2_cudnn_sdpa_c3 python tests/ncu_cudnn_target.py > gpurun_out/ncu_cudnn.log 2>&1
tail -3 gpurun_out/ncu_cudnn.log; ls -la gpurun_out/r2_cudnn_sdpa_c3.ncu-rep

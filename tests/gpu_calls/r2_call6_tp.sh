set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${N:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/symm_probe.py > gpurun_out/r2_symm_probe_n$N.log 2>&1; echo rc=$?
grep "probe\|FAIL" gpurun_out/r2_symm_probe_n$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tests/tp_tune.py > gpurun_out/r2_tp_tune_n$N.log 2>&1; echo rc=$?
grep "shape" gpurun_out/r2_tp_tune_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tests/multi_gpu_check.py > gpurun_out/r2_mgc_n$N.log 2>&1; echo rc=$?
grep "rank 0\|FAIL" gpurun_out/r2_mgc_n$N.log

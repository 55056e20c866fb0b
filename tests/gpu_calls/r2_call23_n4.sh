set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${N:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tests/multi_gpu_check.py > gpurun_out/r2_mgc_n$N.log 2>&1; echo rc=$?
grep "rank 0\|FAIL\|PASS\|ok" gpurun_out/r2_mgc_n$N.log | tail -12
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?
tail -c 600 gpurun_out/r2_bench_n$N.err
tail -c 6000 gpurun_out/r2_bench_n$N.json

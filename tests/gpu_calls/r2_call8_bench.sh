set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${N:-2}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?
grep -i "parity check failed\|Error" gpurun_out/r2_bench_n$N.err | head -3 | cut -c1-600
tail -c 300 gpurun_out/r2_bench_n$N.json

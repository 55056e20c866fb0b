set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
N=${N:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/symm_probe.py > gpurun_out/r2_symm_probe_n$N.log 2>&1; echo rc=$?
tail -30 gpurun_out/r2_symm_probe_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?
tail -c 3000 gpurun_out/r2_bench_n$N.err
cat gpurun_out/r2_bench_n$N.json | tail -c 6000

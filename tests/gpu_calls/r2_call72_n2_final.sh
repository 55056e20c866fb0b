#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_final_multi_gpu_tests.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/r2_final_multi_gpu_tests.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err; echo bench rc=$?
tail -c 300 gpurun_out/r2_final_bench_n$N.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','parity','gpu_launches')})
print(d.get('ring_c5'))
PY

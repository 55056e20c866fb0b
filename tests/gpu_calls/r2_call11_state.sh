set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest4.log 2>&1
tail -6 gpurun_out/r2_pytest4.log
timeout 900 python bench.py --steps 50 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?
tail -c 600 gpurun_out/r2_bench_n1.err
tail -c 6000 gpurun_out/r2_bench_n1.json

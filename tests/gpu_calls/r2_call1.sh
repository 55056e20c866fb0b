set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest1.log 2>&1
tail -5 gpurun_out/r2_pytest1.log
python bench.py --steps 20 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo rc=$?
tail -c 1500 gpurun_out/r2_bench1.err
for g in 2048 4096 16384 32768; do python bench.py --steps 10 --no-secondary --no-ring --no-cpu-baseline --group-rows $g > gpurun_out/r2_bench_g$g.json 2> gpurun_out/r2_bench_g$g.err; done
python tests/ncu_target.py mlp 2 > gpurun_out/ncu_plain_mlp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_act -c 4 -o gpurun_out/prof_mlp_r2a python tests/ncu_target.py mlp 2 > gpurun_out/ncu_mlp_r2a.log 2>&1
ls -la gpurun_out | tail -20

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/ncu_target.py fa 2 > gpurun_out/ncu_plain_fa.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 1 -c 1 -o gpurun_out/prof_fa_r2b python tests/ncu_target.py fa 2 > gpurun_out/ncu_fa.log 2>&1
ls -la gpurun_out/*.ncu-rep

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "layernorm or ln" > gpurun_out/r2_pytest_ln.log 2>&1; tail -3 gpurun_out/r2_pytest_ln.log
timeout 300 python tests/ln_probe.py > gpurun_out/r2_ln_probe2.log 2>&1; cat gpurun_out/r2_ln_probe2.log
SECS=3 timeout 300 python tests/fa_power_probe.py > gpurun_out/r2_fa_power.log 2>&1; cat gpurun_out/r2_fa_power.log

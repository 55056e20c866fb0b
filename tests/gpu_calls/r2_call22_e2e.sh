cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -3; grep -E "AnonHugePages|HugePages_Total|Hugepagesize" /proc/meminfo; cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 120 python tests/e2e_probe.py 2>&1 | head -3
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
free -g | head -3
timeout 120 python tests/e2e_probe.py 2>&1 | head -3

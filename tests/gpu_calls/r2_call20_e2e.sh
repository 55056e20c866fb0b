cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/e2e_bisect.py 2>&1 | tail -8
nvidia-smi topo -m 2>/dev/null | head -8; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv

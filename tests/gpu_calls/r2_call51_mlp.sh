cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "mlp or linear or golden or fullsize or modules" 2>&1 | tail -3

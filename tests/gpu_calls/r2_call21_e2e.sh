cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for s in 50 10; do timeout 300 python bench.py --steps $s --no-secondary --no-ring --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps',d['steps'],'ms/step',round(d['ms_per_step'],3),'e2e ms', round(d['e2e']['ms_per_step'],2))"; done
timeout 300 python bench.py --steps 50 --no-ring 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full-ish steps',d['steps'],'ms/step',round(d['ms_per_step'],3),'e2e ms', round(d['e2e']['ms_per_step'],2))"

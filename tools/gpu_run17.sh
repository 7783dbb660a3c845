cd $GRAFT_REPO_ROOT
python tools/perf_static_eval.py 4194304 1:8 | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['total_ms'], d['Mcfg_per_s'], d['kernel_ms'], d['checksum'])
"

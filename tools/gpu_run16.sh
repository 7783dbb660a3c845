cd $GRAFT_REPO_ROOT
python tools/perf_static_eval.py 4194304 1:8 | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['total_ms'], d['Mcfg_per_s'], d['kernel_ms'], d['checksum'])
"
(timeout 600 python -m pytest tests/test_gpu_headline.py -q -x -k "position_curve or headline_static_parity" 2>&1 | tail -2)

cd $GRAFT_REPO_ROOT
python tools/perf_static3.py 1048576 1:8 1:33 2>&1 | grep -v "one_step_option\": 1" | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['case'], d['ms'], d['Mcfg_per_s'], 'fb', d['fallbacks'], d['fallback_reasons'])
"

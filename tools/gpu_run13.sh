cd $GRAFT_REPO_ROOT
python tools/perf_static3.py 1048576 1:8 1:33 2>&1 | grep -v "one_step_option\": 1" | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['case'], d['ms'], d['Mcfg_per_s'], 'fb', d['fallbacks'], d['fallback_reasons'], 'conv', d['converged'], d['nfev'], d['info_hist'])
"
(timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_static.py -q -x -s 2>&1 | grep "HEADLINE_PARITY\|ONE_STEP_RULE\|passed\|failed\|Error" | cut -c1-1100)

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in default $(ls build/variants/ 2>/dev/null); do
  if [ "$v" == "default" ]; then unset CTRB200_LIB; else export CTRB200_LIB=$GRAFT_REPO_ROOT/build/variants/$v; fi
  echo "== $v"
  python tools/perf_static3.py 1048576 1:8 2:8 2>&1 | grep -v "one_step_option\": 1" | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['case'], d['ms'], d['Mcfg_per_s'], 'fb', d['fallbacks'], 'conv', d['converged'], 'nfev', d['nfev']['mean'], d.get('md5'))
"
  true 2>&1 | grep -v "one_step_option\": 1" | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['case'], d['ms'], d['Mcfg_per_s'], 'fb', d['fallbacks'], 'conv', d['converged'], 'nfev', d['nfev']['mean'], d.get('md5'))
"
done

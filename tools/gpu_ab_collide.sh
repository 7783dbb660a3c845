cd $GRAFT_REPO_ROOT
for v in default $(ls build/variants/); do
  if [ "$v" == "default" ]; then unset CTRB200_LIB; else export CTRB200_LIB=$GRAFT_REPO_ROOT/build/variants/$v; fi
  echo "== $v"
  python tools/perf_static_eval.py 1048576 1:8 1:33 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['case'], 'chain_eval', d['kernel_ms']['chain_eval'], 'total', d['total_ms'], 'collide', d['collide'], 'checksum', d['checksum'])
"
done

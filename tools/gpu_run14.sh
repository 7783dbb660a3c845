cd $GRAFT_REPO_ROOT
for v in default $(ls build/variants/ 2>/dev/null); do
  if [ "$v" == "default" ]; then unset CTRB200_LIB; else export CTRB200_LIB=$GRAFT_REPO_ROOT/build/variants/$v; fi
  echo "== $v"; python tools/perf_static_eval.py 1048576 1:8 | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['total_ms'], d['kernel_ms']['chain_eval'], d['work']['exact_tests'], d['checksum'])
"
  python tools/perf_static_eval.py 262144 1:33 | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('uniform', d['total_ms'], d['kernel_ms']['chain_eval'], d['work']['exact_tests'], d['checksum'])
"
  python tools/perf_eval.py 2>&1 | tail -4
done

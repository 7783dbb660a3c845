cd $GRAFT_REPO_ROOT
for v in default $(ls build/variants/ 2>/dev/null); do
  if [ "$v" == "default" ]; then unset CTRB200_LIB; else export CTRB200_LIB=$GRAFT_REPO_ROOT/build/variants/$v; fi
  echo "== $v"; python tools/perf_eta_ftl.py
done
unset CTRB200_LIB
(timeout 600 python -m pytest tests/test_gpu_kinematic.py tests/test_gpu_solve.py -q -x 2>&1 | tail -3)

cd $GRAFT_REPO_ROOT
bash tools/gpu_perf_ab.sh
for v in default lib_res4_rebase.so; do
  if [ "$v" == "default" ]; then unset CTRB200_LIB; else export CTRB200_LIB=$GRAFT_REPO_ROOT/build/variants/$v; fi
  echo "== tests $v"
  (timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_static.py -q -x -s 2>&1 | grep "HEADLINE_PARITY\|ONE_STEP_RULE\|passed\|failed\|Error" | cut -c1-900)
done

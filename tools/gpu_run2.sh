set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/diag_violators.py > gpurun_out/diag_violators.log 2>&1
cat gpurun_out/diag_violators.log | tail -30
(timeout 900 python -m pytest tests/test_gpu_headline.py -q -s 2>&1 | tail -60) > gpurun_out/pytest_headline_b.log 2>&1
grep -n "HEADLINE_PARITY\|ONE_STEP_QR\|passed\|failed\|Error" gpurun_out/pytest_headline_b.log | cut -c1-3000
# ncu: the fast kernel on the headline mix (30000 configurations)
ncu --set full --clock-control none --import-source on -k regex:static_fast_kernel -c 1 -s 1 -o gpurun_out/r2_fast_a -f python tools/ncu_static.py 30000 1 8 > gpurun_out/ncu_r2_fast_a.log 2>&1
tail -3 gpurun_out/ncu_r2_fast_a.log
ls -la gpurun_out/*.ncu-rep

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_headline.py -q -s 2>&1 | tail -80) > gpurun_out/pytest_headline_c.log 2>&1
grep -n "HEADLINE_PARITY\|ONE_STEP_QR\|passed\|failed\|Error" gpurun_out/pytest_headline_c.log | cut -c1-3500
(timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_headline.py 2>&1 | tail -15) > gpurun_out/pytest_gpu_c.log 2>&1
tail -3 gpurun_out/pytest_gpu_c.log
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_c.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_c.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['fast_kernel_ms'], d['compute_roofline']['collision_kernel_ms'])
PY

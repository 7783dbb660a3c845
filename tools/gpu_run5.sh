set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/pytest_gpu_e.log 2>&1
tail -4 gpurun_out/pytest_gpu_e.log
python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_e.log 2> gpurun_out/bench_e.err
tail -c 400 gpurun_out/bench_e.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_e.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'], d['best_node'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

#!/bin/bash
# The round-end sequence on a GPU box (run through gpurun from the repo root): GPU tests, smoke, both bench arms.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_verify.sh'
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1
(time python bench.py --steps 20 --warmup 3) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 300 gpurun_out/bench_1gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_1gpu.json') if l.startswith('{')][-1])
r = json.loads([l for l in open('gpurun_out/bench_reference.json') if l.startswith('{')][-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "kernel_ms", d['kernel_ms'], "launches", d['gpu_launches'],
      "roofline frac", d['roofline']['frac'], "traffic", d['roofline']['traffic'], "reference", r['value'], "e2e ratio", d['e2e']['value'] / r['value'])
PY

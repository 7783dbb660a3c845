cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time python bench.py --steps 20 --warmup 3) > gpurun_out/bench_r2_1gpu.json 2> gpurun_out/bench_r2_1gpu.err
tail -c 300 gpurun_out/bench_r2_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_1gpu.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('gpurun_out/bench_r2_reference.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'], 'ref', r['value'])
for k,v in d.get('extra',{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('configs_per_s','ms','collide_fraction','converged_fraction','hbm_gbs','frac_of_hbm_peak','hbm_write_gbs')})
print(d['cpu_baseline']['sample'][:1400])
PY

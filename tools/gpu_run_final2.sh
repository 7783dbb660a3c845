cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python bench.py --steps 20 --warmup 3) > gpurun_out/bench_r2_1gpu.json 2> gpurun_out/bench_r2_1gpu.err
tail -c 300 gpurun_out/bench_r2_1gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_1gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'])
for k,v in d.get('extra',{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('configs_per_s','ms','collide_fraction','converged_fraction','hbm_gbs','frac_of_hbm_peak','hbm_write_gbs')})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_static.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'static_fast_kernel|static_solve_kernel|static_curves_pos_kernel|chain_eval_group_kernel' -s 4 -c 4 -o gpurun_out/r2_bench_scale -f python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/ncu_bench_scale.log 2>&1
tail -2 gpurun_out/ncu_bench_scale.log | cut -c1-200

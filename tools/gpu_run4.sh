set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python bench.py --steps 10 --warmup 3) > gpurun_out/bench_d.log 2> gpurun_out/bench_d.err
tail -c 600 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_d.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['gpu_launches'])
print(json.dumps(d['roofline'])[:1500])
print(json.dumps(d.get('cpu_baseline'))[:2500])
for k,v in d.get('extra',{}).items(): print(k, json.dumps(v)[:400])
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_d_ref.log 2>&1
tail -c 1200 gpurun_out/bench_d_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_static.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_d.log 2>&1
tail -2 gpurun_out/ncu_launches_d.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:'static_fast_kernel|static_solve_kernel|static_curves_pos_kernel|chain_eval_group_kernel' -s 4 -c 4 -o gpurun_out/r2_bench_scale -f python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/ncu_bench_scale.log 2>&1
tail -3 gpurun_out/ncu_bench_scale.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# Launch list and one `ncu --set full` capture of bench.py's timed step at bench scale (10^7 configurations per step);
# run on the GPU box AFTER the same commands have exited 0 without ncu:
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile_bench_scale.sh'
# then here: python tools/ncu_bench_scale.py gpurun_out/r2_bench_scale.ncu-rep c3_static_colon_straight_shallow 10000000 > profiles/r2_ncu_bench_scale.json
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --timed-only > gpurun_out/profile_plain.log 2>&1 || { tail -5 gpurun_out/profile_plain.log; exit 1; }
tail -c 400 gpurun_out/profile_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_static_1e7.csv \
    python bench.py --steps 2 --warmup 3 --timed-only > gpurun_out/profile_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'static_fast_kernel|static_solve_kernel|static_curves_pos_kernel|chain_eval_group_kernel' -s 4 -c 4 -f \
    -o gpurun_out/r2_bench_scale python bench.py --steps 1 --warmup 1 --timed-only > gpurun_out/profile_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/r2_bench_scale.ncu-rep gpurun_out/r2_launches_bench_static_1e7.csv

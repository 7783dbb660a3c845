cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:kin_eta_ftl_kernel -s 2 -c 1 -o gpurun_out/r2_eta_ftl -f python tools/perf_eta_ftl.py > gpurun_out/ncu_eta_ftl.log 2>&1
tail -2 gpurun_out/ncu_eta_ftl.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:static_fast_kernel -s 1 -c 1 -o gpurun_out/r2_fast_uniform -f python tools/perf_static3.py 131072 1:33 > gpurun_out/ncu_fast_uniform.log 2>&1
tail -2 gpurun_out/ncu_fast_uniform.log | cut -c1-200
ls -la gpurun_out/r2_*.ncu-rep

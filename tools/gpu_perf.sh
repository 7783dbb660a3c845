cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/perf_static3.py 1048576 1:8 2:8 2>&1 | grep -v "one_step_option\": 1" | tee gpurun_out/perf_static3_last.log
python tools/perf_static3.py 262144 1:33 2>&1 | grep -v "one_step_option\": 1" | tee -a gpurun_out/perf_static3_last.log
(timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_static.py -q -x 2>&1 | tail -5)

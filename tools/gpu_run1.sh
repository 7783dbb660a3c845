set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
(time python tools/perf_static3.py 1048576 1:8 1:33 2:8) > gpurun_out/perf_static3_a.log 2>&1
tail -8 gpurun_out/perf_static3_a.log
(time timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_headline.py 2>&1 | tail -30) > gpurun_out/pytest_gpu_a.log 2>&1
tail -5 gpurun_out/pytest_gpu_a.log
(time timeout 900 python -m pytest tests/test_gpu_headline.py -q -s 2>&1 | tail -60) > gpurun_out/pytest_headline_a.log 2>&1
tail -30 gpurun_out/pytest_headline_a.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/bench_a.log 2>&1
tail -c 3000 gpurun_out/bench_a.log

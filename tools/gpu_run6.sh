cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/run_c5.py 2>&1 | tail -1 | tee gpurun_out/c5_r2_1gpu.json
C5_SPECULATE=3 python tools/run_c5.py 2>&1 | tail -1 | tee gpurun_out/c5_r2_1gpu_spec3.json

cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_headline.py -q -x -k "host_memory_pipeline" 2>&1 | tail -3)

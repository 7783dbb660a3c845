cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_static.py -q -x -s 2>&1 | grep "gpu vs oracle\|passed\|failed\|Error\|assert" | cut -c1-700)

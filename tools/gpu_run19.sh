cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_kinematic.py tests/test_abi.py -m gpu -q -x 2>&1 | tail -5)

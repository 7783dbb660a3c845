cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^HEADLINE_PARITY" | tail -12 | cut -c1-600)

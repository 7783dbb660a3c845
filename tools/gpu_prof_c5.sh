cd $GRAFT_REPO_ROOT
C5_EVALS=20 python -c "
import cProfile, pstats, runpy, sys, io
sys.argv=['tools/run_c5.py']
pr=cProfile.Profile(); pr.enable()
try:
    runpy.run_path('tools/run_c5.py', run_name='__main__')
finally:
    pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
" 2>&1 | tail -60

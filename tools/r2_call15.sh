#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -c "
import cProfile, pstats, io, sys, runpy
sys.argv = ['examples/scordelis_lo.py', '256', '1.0']
pr = cProfile.Profile(); pr.enable()
try:
    runpy.run_path('examples/scordelis_lo.py', run_name='__main__')
finally:
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(70); print(s.getvalue()[:14000])
" > gpurun_out/r2c15_shell_profile.txt 2>&1
grep -v "^Solver" gpurun_out/r2c15_shell_profile.txt | head -95

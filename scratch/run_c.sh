python scratch/fit_bench.py 2>&1 | tail -3

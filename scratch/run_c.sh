SVMATH=bf16 python scratch/sv_bench.py 0 1 2 8 3 2>&1 | grep debug

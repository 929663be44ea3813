SVMATH=bf16 DNB_SV_STAGE=1 python scratch/sv_bench.py 0 2>&1 | grep debug
SVMATH=bf16 python scratch/sv_bench.py 0 2>&1 | grep debug

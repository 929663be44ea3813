timeout 900 python -m pytest tests/test_gpu_layers.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -5
SVMATH=bf16 python scratch/sv_bench.py 0 2>&1 | grep debug
SVMATH=tf32 python scratch/sv_bench.py 0 2>&1 | grep debug
SVMATH=bf16 DNB_SV_PROF=1 python scratch/sv_bench.py 0 2>&1 | grep -E "wg prof\]" | tail -1

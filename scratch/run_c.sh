timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_nets.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -4
SVMATH=bf16 python scratch/sv_bench.py 0 2>&1 | grep debug
SVMATH=tf32 python scratch/sv_bench.py 0 2>&1 | grep debug

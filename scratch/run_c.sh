timeout 900 python -m pytest tests/test_gpu_layers.py -m gpu -x -q -k "loss" 2>&1 | tail -4

timeout 900 python -m pytest tests/test_gpu_layers.py -m gpu -x -q -k "fixtures" 2>&1 | tail -4

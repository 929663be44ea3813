timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_nets.py -m gpu -x -q -k "head or ense or mlp or cnn" 2>&1 | tail -3

timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_nets.py -m gpu -x -q -k "conv or res_conv" 2>&1 | tail -3

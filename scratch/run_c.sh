timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_layers.py -m gpu -x -q -k "fit or evaluate or iris or reg or save_load" 2>&1 | tail -3

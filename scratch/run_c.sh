timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -x -q -k "evaluate or fit or save_load" 2>&1 | tail -6

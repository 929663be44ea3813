timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_dp.py -m gpu -x -q -k "evaluate or fit or save_load or dp or two_rank" 2>&1 | tail -8

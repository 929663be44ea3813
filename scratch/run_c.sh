export DNB_TC_THIN=1
python scratch/tw_bench.py 0 1 7 2>&1 | grep debug
timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -x -q -k "big_conv1" 2>&1 | tail -3

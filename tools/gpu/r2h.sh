bash tools/gpu/r2f.sh
timeout 600 python -m pytest tests/test_gpu_aggdet.py tests/test_detector_set.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15

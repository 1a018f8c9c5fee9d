set -x
timeout 1200 python -m pytest tests/test_gpu_vertvel.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -25

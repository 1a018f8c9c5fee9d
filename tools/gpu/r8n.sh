set -x
timeout 600 python -m pytest tests/test_gpu_host_cpp.py -q -m gpu 2>&1 | tail -12

set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_cpp.py -q -m gpu -k "analytic or restarts" 2>&1 | tail -5

set -x
timeout 900 python -m pytest tests/test_gpu_diag.py -q -m gpu 2>&1 | tail -30
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -q -m gpu -k "neverworld2 or pi_cavity or config" 2>&1 | tail -8

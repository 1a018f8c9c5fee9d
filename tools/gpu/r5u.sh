set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -x -q -m gpu 2>&1 | tail -3
python tools/exp_variants.py --steps 10 "" ""

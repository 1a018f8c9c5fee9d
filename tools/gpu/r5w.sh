set -x
timeout 900 python -m pytest tests/test_gpu_gradients.py tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -3
python tools/exp_variants.py --steps 10 --null-grad "ADV_FUSE_GRAD=1" "ADV_FUSE_GRAD=0"

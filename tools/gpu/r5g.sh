set -x
timeout 1200 python -m pytest tests/test_gpu_gradients.py tests/test_gpu_multirank.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -12
python tools/exp_variants.py --steps 10 "" > gpurun_out/r5g_variants.jsonl 2> gpurun_out/r5g_variants.err
python tools/exp_variants.py --steps 10 --null-grad "ADV_FUSE_GRAD=1" "ADV_FUSE_GRAD=0" >> gpurun_out/r5g_variants.jsonl 2>> gpurun_out/r5g_variants.err
cat gpurun_out/r5g_variants.jsonl; tail -3 gpurun_out/r5g_variants.err

set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_sizes.py -x -q -m gpu 2>&1 | tail -8
python tools/exp_variants.py --steps 10 "" "ADV_K3_ASYNC=0 ADV_K2_ASYNC=0" "ADV_K3_ASYNC=1 ADV_K2_ASYNC=0" "ADV_K3_ASYNC=0 ADV_K2_ASYNC=1" > gpurun_out/r5i_variants.jsonl 2> gpurun_out/r5i_variants.err
cat gpurun_out/r5i_variants.jsonl; tail -3 gpurun_out/r5i_variants.err

set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -x -q -m gpu 2>&1 | tail -15
python tools/exp_variants.py --steps 10 "" "ADV_CTA_THREADS=224 ADV_E1_THREADS=224" "ADV_CTA_THREADS=320 ADV_E1_THREADS=224" "ADV_CTA_THREADS=448 ADV_E1_THREADS=224" "ADV_CTA_THREADS=224 ADV_E1_THREADS=448" "ADV_CTA_THREADS=256 ADV_E1_THREADS=256" "ADV_CTA_THREADS=512 ADV_E1_THREADS=512" "ADV_CTA_THREADS=672 ADV_E1_THREADS=448" > gpurun_out/r4d_variants.jsonl 2> gpurun_out/r4d_variants.err
cat gpurun_out/r4d_variants.jsonl; tail -3 gpurun_out/r4d_variants.err

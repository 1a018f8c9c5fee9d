set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_cavity.py -x -q -m gpu 2>&1 | tail -3
python tools/exp_variants.py --steps 10 "" "ADV_CTA_N1=224" "ADV_CTA_N1=256" "ADV_CTA_N1=384" "ADV_CTA_K3=384" "ADV_CTA_K3=288" "ADV_CTA_K2=256" "" > gpurun_out/r8e_variants.jsonl 2> gpurun_out/r8e_variants.err
cat gpurun_out/r8e_variants.jsonl; tail -3 gpurun_out/r8e_variants.err

set -x
python tools/exp_variants.py --steps 10 "" "ADV_CTA_THREADS=288" "ADV_CTA_THREADS=320" "ADV_CTA_THREADS=384" "ADV_CTA_THREADS=448" "ADV_CTA_THREADS=192" > gpurun_out/r8d_variants.jsonl 2> gpurun_out/r8d_variants.err
cat gpurun_out/r8d_variants.jsonl; tail -3 gpurun_out/r8d_variants.err

set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/exp_variants.py --steps 10 "" "ADV_CTA_THREADS=192" "ADV_CTA_THREADS=256" "ADV_G_K2=3" "ADV_G_K3=3" "ADV_PF=100" "ADV_PF=400" > gpurun_out/r4f_variants.jsonl 2> gpurun_out/r4f_variants.err
cat gpurun_out/r4f_variants.jsonl

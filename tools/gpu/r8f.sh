set -x
python tools/exp_variants.py --steps 10 "" "ADV_G_LO=2" "ADV_G_LO=6" "ADV_CTA_N1=320" "ADV_CTA_N1=352" "ADV_CTA_N1=288 ADV_G_LO=2" "ADV_PF=100" "ADV_PF=400" "ADV_PF=800" "ADV_CTA_K2=448" "ADV_CTA_K3=160" "" > gpurun_out/r8f_variants.jsonl 2> gpurun_out/r8f_variants.err
cat gpurun_out/r8f_variants.jsonl; tail -3 gpurun_out/r8f_variants.err

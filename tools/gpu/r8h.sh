set -x
python tools/exp_variants.py --steps 10 "" "ADV_G_K2=1" "ADV_G_K3=1" "ADV_G_K2=1 ADV_CTA_K2=448" > gpurun_out/r8h_variants.jsonl 2> gpurun_out/r8h_variants.err
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_K2_REGS_64.so python tools/exp_variants.py --steps 10 "ADV_CTA_K2=256 ADV_G_K2=1" "ADV_CTA_K2=512 ADV_G_K2=1" "ADV_CTA_K2=128 ADV_G_K2=1" 2>> gpurun_out/r8h_variants.err | sed "s/\"variant\": \"/\"variant\": \"k2r64 /" >> gpurun_out/r8h_variants.jsonl
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_K2_REGS_56.so python tools/exp_variants.py --steps 10 "ADV_CTA_K2=288 ADV_G_K2=1" "ADV_CTA_K2=384 ADV_G_K2=1" "ADV_CTA_K2=224 ADV_G_K2=1" "ADV_CTA_K2=288" 2>> gpurun_out/r8h_variants.err | sed "s/\"variant\": \"/\"variant\": \"k2r56 /" >> gpurun_out/r8h_variants.jsonl
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_K2_REGS_48.so python tools/exp_variants.py --steps 10 "ADV_CTA_K2=224 ADV_G_K2=1" "ADV_CTA_K2=448 ADV_G_K2=1" 2>> gpurun_out/r8h_variants.err | sed "s/\"variant\": \"/\"variant\": \"k2r48 /" >> gpurun_out/r8h_variants.jsonl
cat gpurun_out/r8h_variants.jsonl; tail -3 gpurun_out/r8h_variants.err

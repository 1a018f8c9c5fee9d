set -x
python tools/exp_variants.py --steps 10 "" "" > gpurun_out/r8g_variants.jsonl 2> gpurun_out/r8g_variants.err
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_K2_REGS_64.so python tools/exp_variants.py --steps 10 "" "ADV_CTA_K2=256" "ADV_CTA_K2=256 ADV_G_K2=1" "ADV_CTA_K2=512" "ADV_CTA_K2=1024" 2>> gpurun_out/r8g_variants.err | sed "s/\"variant\": \"/\"variant\": \"k2r64 /" >> gpurun_out/r8g_variants.jsonl
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_N1_REGS_48.so python tools/exp_variants.py --steps 10 "" "ADV_CTA_N1=224" "ADV_CTA_N1=448" "ADV_CTA_N1=224 ADV_G_LO=2" 2>> gpurun_out/r8g_variants.err | sed "s/\"variant\": \"/\"variant\": \"n1r48 /" >> gpurun_out/r8g_variants.jsonl
cat gpurun_out/r8g_variants.jsonl; tail -3 gpurun_out/r8g_variants.err

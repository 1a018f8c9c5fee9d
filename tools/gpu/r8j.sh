set -x
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_N1_REGS_64.so python tools/exp_variants.py --steps 10 "" "ADV_CTA_N1=256" "ADV_CTA_N1=512" "ADV_CTA_N1=256 ADV_G_LO=6" "ADV_CTA_N1=256 ADV_G_LO=2" 2> gpurun_out/r8j_variants.err | sed "s/\"variant\": \"/\"variant\": \"n1r64 /" > gpurun_out/r8j_variants.jsonl
cat gpurun_out/r8j_variants.jsonl; tail -3 gpurun_out/r8j_variants.err

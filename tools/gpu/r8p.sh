set -x
python tools/exp_variants.py --steps 10 "" "ADV_CTA_K3=288" "" > gpurun_out/r8p_variants.jsonl 2> gpurun_out/r8p_variants.err
for v in pfown7 pfown1 pfown4 pfown0; do FESOM_ADV_LIB=$PWD/build_var/lib_$v.so python tools/exp_variants.py --steps 10 "" 2>> gpurun_out/r8p_variants.err | sed "s/\"variant\": \"/\"variant\": \"$v /" >> gpurun_out/r8p_variants.jsonl; done
cat gpurun_out/r8p_variants.jsonl; tail -2 gpurun_out/r8p_variants.err

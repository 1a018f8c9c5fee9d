set -x
python tools/exp_variants.py --steps 10 "" > gpurun_out/r5q_variants.jsonl 2> gpurun_out/r5q_variants.err
FESOM_ADV_LIB=$PWD/build_var/lib_pfl1.so python tools/exp_variants.py --steps 10 "" "ADV_E1_NG=16" "ADV_E1_NG=4" 2>> gpurun_out/r5q_variants.err | sed "s/\"variant\": \"/\"variant\": \"pfl1 /" >> gpurun_out/r5q_variants.jsonl
cat gpurun_out/r5q_variants.jsonl; tail -3 gpurun_out/r5q_variants.err

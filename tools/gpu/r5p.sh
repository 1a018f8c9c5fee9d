set -x
python tools/exp_variants.py --steps 10 "" > gpurun_out/r5p_variants.jsonl 2> gpurun_out/r5p_variants.err
for v in h200 h1000 h5000; do FESOM_ADV_LIB=$PWD/build_var/lib_$v.so python tools/exp_variants.py --steps 10 "" 2>> gpurun_out/r5p_variants.err | sed "s/\"variant\": \"/\"variant\": \"$v /" >> gpurun_out/r5p_variants.jsonl; done
cat gpurun_out/r5p_variants.jsonl; tail -3 gpurun_out/r5p_variants.err

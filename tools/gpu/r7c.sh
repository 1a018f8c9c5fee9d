set -x
python tools/exp_variants.py --steps 10 "" "ADV_PIPE_K3=2" "ADV_PIPE_K3=3" "ADV_PIPE_K3=4" > gpurun_out/r7c_variants.jsonl 2> gpurun_out/r7c_variants.err
FESOM_ADV_LIB=$PWD/build_var/lib_k3reg72.so python tools/exp_variants.py --steps 10 "ADV_PIPE_K3=3" "ADV_PIPE_K3=4" 2>> gpurun_out/r7c_variants.err | sed "s/\"variant\": \"/\"variant\": \"k3reg72 /" >> gpurun_out/r7c_variants.jsonl
cat gpurun_out/r7c_variants.jsonl; tail -3 gpurun_out/r7c_variants.err

set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python tools/exp_variants.py --steps 10 "" > gpurun_out/r5o_variants.jsonl 2> gpurun_out/r5o_variants.err
FESOM_ADV_LIB=$PWD/build_var/lib_ldg.so python tools/exp_variants.py --steps 10 "" 2>> gpurun_out/r5o_variants.err | sed "s/\"variant\": \"/\"variant\": \"ldg /" >> gpurun_out/r5o_variants.jsonl
python tools/exp_variants.py --steps 10 "" >> gpurun_out/r5o_variants.jsonl 2>> gpurun_out/r5o_variants.err
cat gpurun_out/r5o_variants.jsonl; tail -3 gpurun_out/r5o_variants.err

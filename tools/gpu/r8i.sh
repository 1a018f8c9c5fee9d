set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_cavity.py tests/test_gpu_sizes.py -x -q -m gpu 2>&1 | tail -3
python tools/exp_variants.py --steps 10 "" "" > gpurun_out/r8i_variants.jsonl 2> gpurun_out/r8i_variants.err
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_E1P_REGS_56.so python tools/exp_variants.py --steps 10 "" "ADV_E1_D=3" "ADV_E1_NG=4" "ADV_E1_NG=16" 2>> gpurun_out/r8i_variants.err | sed "s/\"variant\": \"/\"variant\": \"e1r56 /" >> gpurun_out/r8i_variants.jsonl
FESOM_ADV_LIB=$PWD/build_var/lib_ADV_E1P_REGS_48.so python tools/exp_variants.py --steps 10 "" "ADV_E1_D=3" 2>> gpurun_out/r8i_variants.err | sed "s/\"variant\": \"/\"variant\": \"e1r48 /" >> gpurun_out/r8i_variants.jsonl
cat gpurun_out/r8i_variants.jsonl; tail -3 gpurun_out/r8i_variants.err

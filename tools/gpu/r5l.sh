set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_sizes.py tests/test_gpu_vertvel.py -x -q -m gpu 2>&1 | tail -6
python tools/exp_variants.py --steps 10 "" "ADV_G_K3=3" "ADV_G_K3=6" "ADV_G_K2=1" > gpurun_out/r5l_variants.jsonl 2> gpurun_out/r5l_variants.err
for v in k3r56 k3r64 k3r48 k2r80k3r56; do FESOM_ADV_LIB=$PWD/build_var/lib_$v.so python tools/exp_variants.py --steps 10 "" "ADV_G_K3=3" 2>> gpurun_out/r5l_variants.err | sed "s/\"variant\": \"/\"variant\": \"$v /" >> gpurun_out/r5l_variants.jsonl; done
cat gpurun_out/r5l_variants.jsonl; tail -3 gpurun_out/r5l_variants.err

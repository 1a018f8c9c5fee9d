set -x
ADV_LP_K3=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config1 or config2 or scheme_combinations" 2>&1 | tail -3
python tools/exp_variants.py --steps 10 "" "ADV_LP_K3=2" "ADV_LP_K3=2 ADV_G_K3=1" "ADV_LP_K3=2 ADV_G_K3=3" "ADV_LP_K3=2 ADV_CTA_THREADS=128" "ADV_LP_K3=2 ADV_CTA_THREADS=160" "ADV_LP_K3=2 ADV_CTA_THREADS=256" "ADV_LP_K3=2 ADV_G_K3=1 ADV_CTA_THREADS=128" > gpurun_out/r7a_variants.jsonl 2> gpurun_out/r7a_variants.err
for v in k3r80 k3r128; do FESOM_ADV_LIB=$PWD/build_var/lib_$v.so python tools/exp_variants.py --steps 10 "ADV_LP_K3=2" "ADV_LP_K3=2 ADV_G_K3=1" "ADV_LP_K3=2 ADV_G_K3=3" 2>> gpurun_out/r7a_variants.err | sed "s/\"variant\": \"/\"variant\": \"$v /" >> gpurun_out/r7a_variants.jsonl; done
cat gpurun_out/r7a_variants.jsonl; tail -3 gpurun_out/r7a_variants.err

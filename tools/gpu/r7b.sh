set -x
timeout 600 python -m pytest tests/test_gpu_vertvel.py -x -q -m gpu -k "zlevel or zstar" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -3
python tools/exp_variants.py --steps 10 "" "ADV_LP_K3=2 ADV_G_K3=1" > gpurun_out/r7b_variants.jsonl 2> gpurun_out/r7b_variants.err
FESOM_ADV_LIB=$PWD/build_var/lib_k3r72.so python tools/exp_variants.py --steps 10 "ADV_LP_K3=2 ADV_G_K3=1" "ADV_LP_K3=2 ADV_G_K3=1 ADV_CTA_THREADS=128" "ADV_LP_K3=2 ADV_G_K3=1 ADV_CTA_THREADS=192"  2>> gpurun_out/r7b_variants.err | sed "s/\"variant\": \"/\"variant\": \"k3r72 /" >> gpurun_out/r7b_variants.jsonl
cat gpurun_out/r7b_variants.jsonl; tail -3 gpurun_out/r7b_variants.err

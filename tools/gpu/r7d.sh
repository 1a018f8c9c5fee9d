set -x
timeout 900 python -m pytest tests/test_gpu_diag.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_gradients.py -x -q -m gpu 2>&1 | tail -3
python tools/exp_variants.py --steps 10 "" "ADV_G_K2=3" > gpurun_out/r7d_variants.jsonl 2> gpurun_out/r7d_variants.err
python tools/exp_variants.py --steps 10 --tra-diag "" 2>> gpurun_out/r7d_variants.err | sed "s/\"variant\": \"/\"variant\": \"tra_diag /" >> gpurun_out/r7d_variants.jsonl
cat gpurun_out/r7d_variants.jsonl; tail -3 gpurun_out/r7d_variants.err

set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -x -q -m gpu 2>&1 | tail -8
( time python bench.py --workload core2 --steps 5 ) > gpurun_out/r5b_core2.json 2> gpurun_out/r5b_core2.err; tail -c 3000 gpurun_out/r5b_core2.json; tail -5 gpurun_out/r5b_core2.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r5b_ref.json 2> gpurun_out/r5b_ref.err; tail -c 1500 gpurun_out/r5b_ref.json; tail -5 gpurun_out/r5b_ref.err
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r5b_full.json 2> gpurun_out/r5b_full.err; tail -c 6000 gpurun_out/r5b_full.json; tail -12 gpurun_out/r5b_full.err
free -g | head -2

set -x
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r5c_full.json 2> gpurun_out/r5c_full.err; tail -c 7000 gpurun_out/r5c_full.json; tail -15 gpurun_out/r5c_full.err
free -g | head -2

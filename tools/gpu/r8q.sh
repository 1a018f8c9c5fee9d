set -x
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cavity.py tests/test_gpu_multirank.py -x -q -m gpu -k "config or cavity_mesh_matches or neverworld2 or use_wsplit or (local_ranks_match and (pi-8 or synth-5))" 2>&1 | tail -2
python bench.py --workload cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/r8q_bench_share.json 2>/dev/null; tail -c 900 gpurun_out/r8q_bench_share.json

set -x
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool memcheck --error-exitcode 9 --log-file gpurun_out/r8l_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_cavity.py tests/test_gpu_multirank.py -x -q -m gpu -k "config1 or config2 or neverworld2 or use_wsplit or (cavity_mesh_matches and FCT) or (local_ranks_match and (pi-2 or synth-5))" 2>&1 | tail -3; tail -2 gpurun_out/r8l_memcheck.log
timeout 600 $CS --tool racecheck --error-exitcode 9 --log-file gpurun_out/r8l_racecheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_cavity.py -x -q -m gpu -k "config2 or use_wsplit or (cavity_mesh_matches and QR4C and FCT) or (neverworld2 and QR4C)" 2>&1 | tail -3; tail -2 gpurun_out/r8l_racecheck.log
timeout 300 $CS --tool synccheck --error-exitcode 9 --log-file gpurun_out/r8l_synccheck.log python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config2 or use_wsplit" 2>&1 | tail -3; tail -2 gpurun_out/r8l_synccheck.log

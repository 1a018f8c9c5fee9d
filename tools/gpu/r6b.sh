set -x
CS=/usr/local/cuda/bin/compute-sanitizer
$CS --tool racecheck --log-file gpurun_out/r6b_rc.log python tools/gpu/dbg_nofct.py plain
timeout 900 $CS --tool racecheck --error-exitcode 9 --log-file gpurun_out/r6b_racecheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_vertvel.py tests/test_gpu_gradients.py -x -q -m gpu -k "config2 or kernel_variants or use_wsplit or (scheme_combinations and QR4C) or vert_vel_ale_matches and pi or null_gradient" 2>&1 | tail -3; tail -1 gpurun_out/r6b_racecheck.log
timeout 900 $CS --tool racecheck --error-exitcode 9 --log-file gpurun_out/r6b_racecheck_mr.log python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "local_ranks_match and (pi-2 or synth-5) or device_gradients and synth" 2>&1 | tail -3; tail -1 gpurun_out/r6b_racecheck_mr.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3

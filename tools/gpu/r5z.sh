set -x
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool racecheck --error-exitcode 9 --log-file gpurun_out/r5z_racecheck.log python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scheme_combinations and QR4C" 2>&1 | tail -40; tail -5 gpurun_out/r5z_racecheck.log

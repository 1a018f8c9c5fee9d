set -x
export ADV_SAN=1
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 9 --log-file gpurun_out/r5d_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_gradients.py -x -q -m gpu -k "config1 or config2 or scheme_combinations or use_wsplit or kernel_variants or odd_tracer or unaligned or null_gradient or reference_call_order" 2>&1 | tail -5
echo "memcheck rc=$?"; tail -5 gpurun_out/r5d_memcheck.log
timeout 1500 $CS --tool memcheck --error-exitcode 9 --log-file gpurun_out/r5d_memcheck_mr.log python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "local_ranks_match and (pi-2 or synth-5) or device_gradients and synth" 2>&1 | tail -5
tail -5 gpurun_out/r5d_memcheck_mr.log
timeout 1500 $CS --tool racecheck --error-exitcode 9 --log-file gpurun_out/r5d_racecheck.log python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config2 or kernel_variants or use_wsplit or (scheme_combinations and FCT and QR4C)" 2>&1 | tail -5
tail -8 gpurun_out/r5d_racecheck.log
timeout 900 $CS --tool synccheck --error-exitcode 9 --log-file gpurun_out/r5d_synccheck.log python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config2 or kernel_variants" 2>&1 | tail -5
tail -5 gpurun_out/r5d_synccheck.log

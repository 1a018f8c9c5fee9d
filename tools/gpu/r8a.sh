set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 --log-file gpurun_out/r8a_memcheck.log python -m pytest tests/test_gpu_diag.py tests/test_gpu_cavity.py tests/test_gpu_vertvel.py tests/test_gpu_parity.py -x -q -m gpu -k "ltra_diag or ldiag_dvd and not True or cavity or zlevel or neverworld2 or config2" 2>&1 | tail -3; tail -2 gpurun_out/r8a_memcheck.log
timeout 600 $CS --tool racecheck --error-exitcode 9 --log-file gpurun_out/r8a_racecheck.log python -m pytest tests/test_gpu_diag.py tests/test_gpu_cavity.py -x -q -m gpu -k "(ltra_diag_matches and False and FCT) or (ldiag_dvd and False and pi) or (cavity_mesh_matches and QR4C)" 2>&1 | tail -3; tail -2 gpurun_out/r8a_racecheck.log
ncu --set full --clock-control none --import-source on -k regex:'k_(edge_flux|node_lo|fct)' -s 16 -c 4 -o gpurun_out/r8a_step -f python tools/exp_variants.py --steps 1 "" > gpurun_out/r8a_ncu.log 2>&1
tail -1 gpurun_out/r8a_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/r8a_launches.csv python bench.py --steps 2 --warmup 3 --no-parity --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r8a_launches.log 2>&1
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r8a_full.json 2> gpurun_out/r8a_full.err; tail -c 1500 gpurun_out/r8a_full.json; grep "bench\]\|real" gpurun_out/r8a_full.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r8a_ref.json 2> gpurun_out/r8a_ref.err; tail -c 400 gpurun_out/r8a_ref.json; grep real gpurun_out/r8a_ref.err
python bench.py --workload cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r8a_bench_share.json 2> gpurun_out/r8a_bench_share.err; tail -c 1500 gpurun_out/r8a_bench_share.json

set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:'k_(edge_flux|node_lo|fct)' -s 16 -c 4 -o gpurun_out/r8k_step -f python tools/exp_variants.py --steps 1 "" > gpurun_out/r8k_ncu.log 2>&1
tail -1 gpurun_out/r8k_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/r8k_launches.csv python bench.py --steps 2 --warmup 3 --no-parity --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r8k_launches.log 2>&1
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r8k_full.json 2> gpurun_out/r8k_full.err; tail -c 1500 gpurun_out/r8k_full.json; grep "bench\]\|real" gpurun_out/r8k_full.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r8k_ref.json 2> gpurun_out/r8k_ref.err; tail -c 400 gpurun_out/r8k_ref.json; grep real gpurun_out/r8k_ref.err
python bench.py --workload cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r8k_bench_share.json 2> gpurun_out/r8k_bench_share.err; tail -c 1500 gpurun_out/r8k_bench_share.json

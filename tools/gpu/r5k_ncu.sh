set -x
# launch list of the bench command (default workload: 3.0M x 70 on one GPU; legs that launch no library kernels switched off)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/r5k_launches.csv python bench.py --steps 2 --warmup 3 --no-parity --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r5k_launches.log 2>&1
tail -1 gpurun_out/r5k_launches.log | cut -c1-400
# full capture of one step of the per-GPU share (613x613x70), contract path
ncu --set full --clock-control none --import-source on -k regex:'k_(edge_flux|node_lo|fct)' -s 16 -c 4 -o gpurun_out/r5k_step -f python tools/exp_variants.py --steps 1 "" > gpurun_out/r5k_ncu.log 2>&1
tail -2 gpurun_out/r5k_ncu.log
python tools/exp_variants.py --steps 10 "" "ADV_PF=100"

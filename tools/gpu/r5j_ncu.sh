set -x
# launch list of the bench command (default workload: 3.0M x 70 on one GPU), legs that launch no library kernels switched off
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5j_launches.csv python bench.py --steps 2 --warmup 3 --no-parity --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r5j_launches.log 2>&1
tail -2 gpurun_out/r5j_launches.log | cut -c1-300
# full capture of one step of the per-GPU share (613x613x70), contract path
ncu --set full --clock-control none --import-source on -k regex:'k_(edge_flux|node_lo|fct)' -s 16 -c 4 -o gpurun_out/r5j_step -f python tools/exp_variants.py --steps 1 "" > gpurun_out/r5j_ncu.log 2>&1
tail -2 gpurun_out/r5j_ncu.log
# full capture of the fused-gradient step (edge_up_dn_grad = NULL): edge kernel + the two producers
ncu --set full --clock-control none --import-source on -k regex:'k_(edge_flux|tracer_gradient|node_mean)' -s 20 -c 5 -o gpurun_out/r5j_fused -f python tools/exp_variants.py --steps 1 --null-grad "" > gpurun_out/r5j_ncu_fused.log 2>&1
tail -2 gpurun_out/r5j_ncu_fused.log

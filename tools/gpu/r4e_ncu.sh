set -x
ncu --set full --clock-control none --import-source on -k regex:'k_(edge_flux|node_lo|fct)' -s 16 -c 4 -o gpurun_out/r4e_compact -f python tools/exp_variants.py --steps 1 "ADV_CTA_THREADS=224 ADV_E1_THREADS=224" > gpurun_out/r4e_ncu.log 2>&1
tail -3 gpurun_out/r4e_ncu.log

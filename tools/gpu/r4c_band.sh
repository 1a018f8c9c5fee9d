set -x
python tools/exp_variants.py --steps 10 "" "ADV_L2HINT=1" "ADV_L2HINT=2" "ADV_BAND=1839 ADV_L2HINT=1" "ADV_BAND=1226 ADV_L2HINT=1" "ADV_BAND=3678 ADV_L2HINT=1" "ADV_BAND=1839 ADV_L2HINT=1 ADV_BAND_STREAMS=1" "ADV_BAND=1226 ADV_L2HINT=1 ADV_BAND_STREAMS=1" "ADV_BAND=3678 ADV_L2HINT=1 ADV_BAND_STREAMS=1" "ADV_BAND=3678 ADV_L2HINT=2 ADV_BAND_STREAMS=1" "ADV_BAND=7356 ADV_L2HINT=1 ADV_BAND_STREAMS=1" > gpurun_out/r4c_variants.jsonl 2> gpurun_out/r4c_variants.err
cat gpurun_out/r4c_variants.jsonl
NCU="ncu --profile-from-start off --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv"
$NCU --log-file gpurun_out/r4c_tr_hint.csv python tools/traffic_step.py ADV_L2HINT=1 > /dev/null 2>&1
$NCU --log-file gpurun_out/r4c_tr_b1839h.csv python tools/traffic_step.py ADV_BAND=1839 ADV_L2HINT=1 > /dev/null 2>&1
$NCU --log-file gpurun_out/r4c_tr_b1226h.csv python tools/traffic_step.py ADV_BAND=1226 ADV_L2HINT=1 > /dev/null 2>&1
$NCU --log-file gpurun_out/r4c_tr_b3678h.csv python tools/traffic_step.py ADV_BAND=3678 ADV_L2HINT=1 > /dev/null 2>&1
python tools/traffic_sum.py gpurun_out/r4c_tr_*.csv

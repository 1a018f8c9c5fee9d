set -x
python tools/exp_variants.py --steps 10 "" "ADV_BAND=1839" "ADV_BAND=3678" "ADV_BAND=7356" "ADV_BAND=14712" "ADV_BAND=1839 ADV_BAND_STREAMS=1" "ADV_BAND=3678 ADV_BAND_STREAMS=1" "ADV_BAND=7356 ADV_BAND_STREAMS=1" "ADV_BAND=14712 ADV_BAND_STREAMS=1" "ADV_BAND=29424 ADV_BAND_STREAMS=1" > gpurun_out/r4b_variants.jsonl 2> gpurun_out/r4b_variants.err
cat gpurun_out/r4b_variants.jsonl
NCU="ncu --profile-from-start off --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv"
$NCU --log-file gpurun_out/r4b_tr_base.csv python tools/traffic_step.py > /dev/null 2>&1
$NCU --log-file gpurun_out/r4b_tr_b1839.csv python tools/traffic_step.py ADV_BAND=1839 > /dev/null 2>&1
$NCU --log-file gpurun_out/r4b_tr_b3678.csv python tools/traffic_step.py ADV_BAND=3678 > /dev/null 2>&1
$NCU --log-file gpurun_out/r4b_tr_b7356.csv python tools/traffic_step.py ADV_BAND=7356 > /dev/null 2>&1
python tools/traffic_sum.py gpurun_out/r4b_tr_*.csv

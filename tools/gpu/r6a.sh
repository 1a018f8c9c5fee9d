set -x
/usr/local/cuda/bin/compute-sanitizer --tool racecheck --log-file gpurun_out/r6a_rc.log python tools/gpu/dbg_nofct.py plain
/usr/local/cuda/bin/compute-sanitizer --tool racecheck --log-file gpurun_out/r6a_rc2.log python tools/gpu/dbg_nofct.py sync

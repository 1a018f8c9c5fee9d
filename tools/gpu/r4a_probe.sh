set -x
for c in gfortran flang flang-new ifort ifx nvfortran pgfortran mpif90 mpifort mpicc mpirun mpiexec f95 f77 nc-config nf-config; do which $c; done
ls /opt/nvidia 2>/dev/null; ls /opt 2>/dev/null
ls /usr/lib/gcc/x86_64-linux-gnu/*/ | head -30
nproc; free -g; ulimit -a | head -20; cat /sys/fs/cgroup/memory.max 2>/dev/null; df -h /dev/shm /tmp | cat
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
lscpu | head -25
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; tail -c 1500 gpurun_out/r4a_bench.json

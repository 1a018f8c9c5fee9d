set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_host_cpp.py tests/test_gpu_diag.py tests/test_gpu_cavity.py -x -q -m gpu -k "nccl or restarts or two_local_ranks" 2>&1 | tail -5
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r8b_bench2.json 2> gpurun_out/r8b_bench2.err; tail -c 3000 gpurun_out/r8b_bench2.json; grep -v "^$" gpurun_out/r8b_bench2.err | tail -8

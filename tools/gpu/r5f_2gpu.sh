set -x
nvidia-smi --query-gpu=index,name --format=csv
nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "nccl" 2>&1 | tail -5
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r5f_bench2.json 2> gpurun_out/r5f_bench2.err; tail -c 5000 gpurun_out/r5f_bench2.json; grep -v "^$" gpurun_out/r5f_bench2.err | tail -15
( time python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r5f_ref2.json 2> gpurun_out/r5f_ref2.err; tail -c 600 gpurun_out/r5f_ref2.json; tail -4 gpurun_out/r5f_ref2.err

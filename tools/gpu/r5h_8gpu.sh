set -x
nvidia-smi --query-gpu=index,name --format=csv | head -3; nproc; free -g | head -2
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "nccl" 2>&1 | tail -5
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r5h_bench8.json 2> gpurun_out/r5h_bench8.err; tail -c 5000 gpurun_out/r5h_bench8.json; grep "bench\]\|real\|Error\|error" gpurun_out/r5h_bench8.err | tail -15
( time python bench.py --impl reference --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r5h_ref8.json 2> gpurun_out/r5h_ref8.err; tail -c 500 gpurun_out/r5h_ref8.json; tail -4 gpurun_out/r5h_ref8.err

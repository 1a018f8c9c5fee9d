#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2d_bench8.json 2> gpurun_out/r2d_bench8.err
cat gpurun_out/r2d_bench8.json; tail -2 gpurun_out/r2d_bench8.err

#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r3b_tests2.log 2>&1
echo "tests exit $?" >> gpurun_out/r3b_tests2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r3b_bench2.json 2> gpurun_out/r3b_bench2.err
tail -3 gpurun_out/r3b_tests2.log; grep '^{"metric' gpurun_out/r3b_bench2.json | cut -c1-330

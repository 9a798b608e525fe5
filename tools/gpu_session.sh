#!/bin/bash
# session 5 (2 GPUs): device timeline of the five-kernel expert-parallel layer
mkdir -p gpurun_out
B2M_TIMELINE=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s_ep2_timeline.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep2_timeline.log
tail -2 gpurun_out/s_ep2_timeline.log | cut -c1-600

#!/bin/bash
# session 6 (2 GPUs): fused route+dispatch kernel, grid sized to the tile list: parity test, timeline, clean bench
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_ep.py -m gpu -q --timeout 380 > gpurun_out/s_pytest_ep.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest_ep.log
tail -15 gpurun_out/s_pytest_ep.log | cut -c1-400
B2M_TIMELINE=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s_ep2_timeline.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep2_timeline.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s_ep2.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep2.log
tail -2 gpurun_out/s_ep2.log | cut -c1-400

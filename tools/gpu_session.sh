#!/bin/bash
# session 4 (2 GPUs): expert-parallel parity test, then the N=2 bench with the five-kernel layer and with the seven-kernel one
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_ep.py -m gpu -q --timeout 380 > gpurun_out/s_pytest_ep.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest_ep.log
tail -30 gpurun_out/s_pytest_ep.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s_ep2_direct.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep2_direct.log
B2M_EP_DIRECT=0 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s_ep2_v1.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep2_v1.log
tail -3 gpurun_out/s_ep2_direct.log | cut -c1-1500; tail -3 gpurun_out/s_ep2_v1.log | cut -c1-1500

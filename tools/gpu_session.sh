#!/bin/bash
# session (2 GPUs): does the fused expert-FFN kernel help the expert-parallel layer?  + the tracer test
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_ep.py -m gpu -q --timeout 280 > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
tail -4 gpurun_out/s_pytest.log | cut -c1-300
B2M_FUSED_FFN=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s_ep2_fused.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep2_fused.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s_ep2.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep2.log
for f in s_ep2_fused s_ep2; do tail -2 gpurun_out/$f.log | cut -c1-260; done

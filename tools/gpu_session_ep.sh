#!/bin/bash
# expert-parallel bench on NG GPUs: timeline run + clean run (+ the seven-kernel sequence for comparison when V1=1)
NG=${NG:-8}
mkdir -p gpurun_out
B2M_TIMELINE=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/s_ep${NG}_timeline.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep${NG}_timeline.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/s_ep${NG}.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep${NG}.log
if [ "${V1:-0}" = "1" ]; then
B2M_EP_DIRECT=0 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/s_ep${NG}_v1.log 2>&1; echo "rc=$?" >> gpurun_out/s_ep${NG}_v1.log
fi
tail -2 gpurun_out/s_ep${NG}.log | cut -c1-300

#!/bin/bash
# one timeline run of the expert-parallel bench on NG GPUs (+ the multi-GPU parity test when PYT=1)
NG=${NG:-2}
mkdir -p gpurun_out
if [ "${PYT:-0}" = "1" ]; then
timeout -k 10 600 python -m pytest tests/test_gpu_ep.py -x -q > gpurun_out/q_pytest${NG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest${NG}.log; tail -3 gpurun_out/q_pytest${NG}.log
fi
B2M_TIMELINE=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/q_ep${NG}.log 2>&1; echo "rc=$?" >> gpurun_out/q_ep${NG}.log
grep -h -o '"ms_per_step": [0-9.]*\|"ep_parity": [a-z]*' gpurun_out/q_ep${NG}.log

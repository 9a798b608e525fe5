#!/bin/bash
# expert-parallel A/B on NG GPUs: multi-GPU parity tests, then bench with and without the I/SMs-row gate/up tiles
NG=${NG:-2}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ep.py -x -q > gpurun_out/ab_pytest${NG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest${NG}.log
port=29620
for mr in 1 0 1 0; do
  port=$((port+1))
  B2M_EP_MROWS=$mr timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/ab_ep${NG}_mr${mr}_$port.log 2>&1; echo "rc=$?" >> gpurun_out/ab_ep${NG}_mr${mr}_$port.log
done
B2M_TIMELINE=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29630 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/ab_ep${NG}_timeline.log 2>&1
tail -3 gpurun_out/ab_pytest${NG}.log
grep -h -o '"ms_per_step": [0-9.]*\|"ep_parity": [a-z]*' gpurun_out/ab_ep${NG}_mr*.log

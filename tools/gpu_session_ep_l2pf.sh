#!/bin/bash
# expert-parallel bench on NG GPUs over L2 prefetch budgets (MB; 0 = no prefetch, no programmatic edges)
NG=${NG:-2}
mkdir -p gpurun_out
if [ "${PYT:-0}" = "1" ]; then
timeout -k 10 300 python -m pytest tests/test_gpu_ep.py -x -q > gpurun_out/l2pf_pytest${NG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l2pf_pytest${NG}.log; tail -3 gpurun_out/l2pf_pytest${NG}.log
fi
port=29650
for mb in ${MBS:-0 48 96}; do
  port=$((port+1))
  B2M_TIMELINE=${TL:-0} B2M_EP_L2PF_MB=$mb timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/l2pf_ep${NG}_mb${mb}.log 2>&1; echo "rc=$?" >> gpurun_out/l2pf_ep${NG}_mb${mb}.log
  echo "mb=$mb $(grep -h -o '"ms_per_step": [0-9.]*\|"ep_parity": [a-z]*' gpurun_out/l2pf_ep${NG}_mb${mb}.log | tr '\n' ' ')"
done

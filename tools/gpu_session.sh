#!/bin/bash
# session 3: look-ahead/activation-aware GPU test, reference engine under forced offload (sequential protocol), config 3 small then full
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_offload.py -m gpu -q --timeout 300 > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 900 python tools/ref_engine_harness.py --mode timing --layers 4 --budget-experts 15 --steps 4 --warmup 1 --compare 0 --protocol sequential --out gpurun_out/ref_timing_offload.json > gpurun_out/s_timing_off.log 2>&1; echo "rc=$?" >> gpurun_out/s_timing_off.log
timeout 600 python bench.py --config offload --layers 8 --steps 6 --warmup 2 --prefill 2048 --ablate > gpurun_out/s_offload_small.log 2>&1; echo "rc=$?" >> gpurun_out/s_offload_small.log
if tail -2 gpurun_out/s_offload_small.log | grep -q '"same_outputs": true'; then
  timeout 1500 python bench.py --config offload --steps 32 --warmup 4 --ablate > gpurun_out/s_offload_full.log 2>&1; echo "rc=$?" >> gpurun_out/s_offload_full.log
fi
tail -4 gpurun_out/s_pytest.log; tail -2 gpurun_out/s_timing_off.log | cut -c1-900; tail -2 gpurun_out/s_offload_small.log | cut -c1-3000; tail -2 gpurun_out/s_offload_full.log | cut -c1-3000

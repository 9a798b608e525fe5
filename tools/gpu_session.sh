#!/bin/bash
# one gpurun call (session 2): GPU test-suite, the reference engine timed beside ours, both bench arms, DeepSeek / DYN_N baselines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 900 python tools/ref_engine_harness.py --mode timing --layers 4 --ratio 0.9 --steps 8 --out gpurun_out/ref_timing_resident.json > gpurun_out/s_timing_res.log 2>&1; echo "rc=$?" >> gpurun_out/s_timing_res.log
timeout 900 python tools/ref_engine_harness.py --mode timing --layers 4 --budget-experts 15 --steps 4 --warmup 1 --compare 0 --out gpurun_out/ref_timing_offload.json > gpurun_out/s_timing_off.log 2>&1; echo "rc=$?" >> gpurun_out/s_timing_off.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/s_bench_ref.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench.log 2>&1; echo "rc=$?" >> gpurun_out/s_bench.log
timeout 600 python tools/bench_configs.py --what deepseek --out gpurun_out/s_deepseek_base.json > gpurun_out/s_deepseek_base.log 2>&1
B2M_DYN_N=1 timeout 600 python tools/bench_configs.py --what deepseek --out gpurun_out/s_deepseek_dyn.json > gpurun_out/s_deepseek_dyn.log 2>&1
tail -4 gpurun_out/s_pytest.log; tail -2 gpurun_out/s_timing_res.log; tail -2 gpurun_out/s_timing_off.log; tail -2 gpurun_out/s_bench_ref.log | cut -c1-600; tail -2 gpurun_out/s_bench.log | cut -c1-300; tail -2 gpurun_out/s_deepseek_base.log; tail -2 gpurun_out/s_deepseek_dyn.log

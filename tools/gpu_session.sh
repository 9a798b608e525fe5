#!/bin/bash
# session: fused expert-FFN kernel -- full GPU test-suite, headline bench with and without it, DeepSeek config
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
tail -5 gpurun_out/s_pytest.log | cut -c1-300
timeout -k 10 400 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s_bench_fused.log 2>&1; echo "rc=$?" >> gpurun_out/s_bench_fused.log
B2M_FUSED_FFN=0 timeout -k 10 400 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s_bench_unfused.log 2>&1; echo "rc=$?" >> gpurun_out/s_bench_unfused.log
timeout -k 10 400 python bench.py --config deepseek --steps 20 --prefill 0 > gpurun_out/s_deepseek_fused.log 2>&1; echo "rc=$?" >> gpurun_out/s_deepseek_fused.log
for f in s_bench_fused s_bench_unfused s_deepseek_fused; do tail -2 gpurun_out/$f.log | cut -c1-260; done

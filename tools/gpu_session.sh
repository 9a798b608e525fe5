#!/bin/bash
# one gpurun call: environment probe, GPU test-suite, the reference engine (policy trace + timing), a short bench
mkdir -p gpurun_out
{ nproc; free -g; df -h . /tmp /var/tmp /dev/shm 2>/dev/null; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/s_env.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 400 python tools/ref_engine_harness.py --mode policy --out gpurun_out/policy_ref_trace.json > gpurun_out/s_policy.log 2>&1; echo "rc=$?" >> gpurun_out/s_policy.log
timeout 700 python tools/ref_engine_harness.py --mode timing --layers 4 --ratio 0.9 --steps 8 --out gpurun_out/ref_timing_resident.json > gpurun_out/s_timing_res.log 2>&1; echo "rc=$?" >> gpurun_out/s_timing_res.log
timeout 700 python tools/ref_engine_harness.py --mode timing --layers 4 --budget-experts 15 --steps 4 --warmup 1 --compare 0 --out gpurun_out/ref_timing_offload.json > gpurun_out/s_timing_off.log 2>&1; echo "rc=$?" >> gpurun_out/s_timing_off.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s_bench.log 2>&1; echo "rc=$?" >> gpurun_out/s_bench.log
tail -5 gpurun_out/s_pytest.log; tail -3 gpurun_out/s_policy.log; tail -2 gpurun_out/s_timing_res.log; tail -2 gpurun_out/s_timing_off.log; tail -2 gpurun_out/s_bench.log

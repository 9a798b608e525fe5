#!/bin/bash
# final single-GPU check of round 2: suite, smoke, both bench arms, ncu launch list, config 3 with the prefetch governor
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_tests.log
tail -4 gpurun_out/r02d_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02d_smoke.log 2>&1; tail -1 gpurun_out/r02d_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02d_bench_reference.log 2>&1; tail -1 gpurun_out/r02d_bench_reference.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.log 2>&1; tail -1 gpurun_out/r02d_bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:grouped_gemm|gate_topk|permute_small|combine_kernel' -c 400 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02d_ncu_launches.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/r02d_launches.csv gpurun_out/r02d_launches.txt > /dev/null 2>&1; head -12 gpurun_out/r02d_launches.txt
timeout 1500 python bench.py --config offload --steps 32 --warmup 4 > gpurun_out/r02d_offload.log 2>&1; tail -1 gpurun_out/r02d_offload.log | cut -c1-200
timeout 1500 python bench.py --config offload --steps 32 --warmup 4 --skew 2.0 > gpurun_out/r02d_offload_skew2.log 2>&1; tail -1 gpurun_out/r02d_offload_skew2.log | cut -c1-200

#!/usr/bin/env bash
# One gpurun call that produces everything a round needs from the GPU box (each gpurun call costs ~1 GPU-minute of
# overhead, so batch):   gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh r02a'
# Outputs under gpurun_out/<tag>_*: test log, smoke, bench line, secondary configs, graph-mode stage microbench,
# ncu launch list (bench step) and one ncu --set full capture of the two grouped GEMM launches of a decode step.
# Copy what should be judged into profiles/ afterwards (tools/ncu_summarize.py makes the text summaries).
set -u
TAG="${1:-check}"
OUT=gpurun_out
mkdir -p "$OUT"
run() { local name="$1"; shift; echo "== $name" ; ( timeout "$1" "${@:2}" ) > "$OUT/${TAG}_${name}.log" 2>&1; echo "   rc=$? (log $OUT/${TAG}_${name}.log)"; }

run tests 900 python -m pytest tests -m gpu -q
tail -3 "$OUT/${TAG}_tests.log"
run smoke 300 python __graft_entry__.py smoke
tail -1 "$OUT/${TAG}_smoke.log"
run bench 400 python bench.py
tail -1 "$OUT/${TAG}_bench.log" > "$OUT/${TAG}_bench.json"
python - "$OUT/${TAG}_bench.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   bench:", d["value"], d["unit"], d["ms_per_step"], "ms/step, e2e", d["e2e"]["value"], "roofline step",
          d["roofline"].get("step", {}).get("frac"), "K3", d["roofline"]["frac"], "clocks", d["clocks"])
except Exception as ex:
    print("   bench line unreadable:", ex)
PY
run configs 600 python tools/bench_configs.py --what "${CONFIGS:-prefill,deepseek}" --out "$OUT/${TAG}_configs.json"
tail -4 "$OUT/${TAG}_configs.log"
run microbench 300 python tools/route_microbench.py
cat "$OUT/${TAG}_microbench.log"
# ncu: launch list of two bench steps (the first ~260 launches are torch weight-init kernels), then a full capture
run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:grouped_gemm|gate_topk|permute_small|combine_kernel|route_' -c 400 --csv \
    --log-file "$OUT/${TAG}_launches.csv" python bench.py --steps 2 --warmup 1
python tools/ncu_summarize.py launches "$OUT/${TAG}_launches.csv" "$OUT/${TAG}_launches.txt" > /dev/null 2>&1 && cat "$OUT/${TAG}_launches.txt"
run ncu_full 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm_tc_kernel -s 4 -c 2 \
    -o "$OUT/${TAG}_k3k4" -f python bench.py --steps 1 --warmup 1
ls -la "$OUT/${TAG}_k3k4.ncu-rep" 2>/dev/null
# optional: the reference's own native engine beside ours (REF_ENGINE=1; needs oracle/_ref/prefetch_op.so and an
# O_DIRECT-capable directory; aborts on any DLOG_FATAL, hence last and in its own process)
if [ "${REF_ENGINE:-0}" = 1 ]; then
  run ref_engine 900 python tools/ref_engine_harness.py --mode timing --layers 4 --tokens 8 --steps 16 --ratio 0.9
  tail -2 "$OUT/${TAG}_ref_engine.log"
fi
# the other BASELINE configs through bench.py's own driver-runnable lines
if [ "${EXTRA:-0}" = 1 ]; then
  run bench_reference 600 python bench.py --impl reference --steps 20 --warmup 5
  tail -1 "$OUT/${TAG}_bench_reference.log" | cut -c1-400
  run deepseek 600 python bench.py --config deepseek --steps 20
  tail -1 "$OUT/${TAG}_deepseek.log" | cut -c1-600
  run offload 1500 python bench.py --config offload --steps 32 --warmup 4 --ablate
  tail -1 "$OUT/${TAG}_offload.log" | cut -c1-400
  run offload_skew2 1500 python bench.py --config offload --steps 32 --warmup 4 --skew 2.0 --ablate
  tail -1 "$OUT/${TAG}_offload_skew2.log" | cut -c1-400
fi
echo "done: $TAG"

#!/usr/bin/env python
"""Drive the reference's OWN native engine (oracle/_ref/prefetch_op.so) on the GPU box  --  TEST / BASELINE INFRASTRUCTURE.

STATUS: written in round 1 after the GPU budget was spent; the module builds and imports in the dev container
(`make -C oracle/ref_build -f Makefile.engine`), but this script has NOT been executed on a GPU yet.

What it does (the "T1" tier of SURVEY §8c / the R-gpu timing of §8d): builds a reduced-L Mixtral-shaped MoE stack, hands
its tensors to the reference engine exactly the way moe_infinity/runtime/model_offload.py does (offload -> register
placeholders -> expert_dispatcher -> set_topology -> register_expert), then per layer performs what
`DistributedExpertExecutor.dispatch_local` + the block's Python combine do (expert_executor.py:32-58, mixtral.py:87-101),
times it, and compares the hidden states with this repository's engine on the same weights and routing.

Gotchas taken from the reference code:
  * exactly one prefetch_handle per process (global singletons, archer_prefetch_handle.cpp:18-28);
  * >= 2 dense stages per GPU or InitializeTopology divides by zero (model_topology.cpp:518,525);
  * the expert_dispatcher is constructed BEFORE set_topology in the real flow (model_offload.py:471-477 vs :606), which is
    what makes its HBM budget ratio x total memory;
  * the offload directory must accept O_DIRECT (tmpfs does not);  any DLOG_FATAL aborts the process -> run as a subprocess;
  * never destroy the engine objects: destructors join threads blocked on condition variables -> os._exit at the end.

  python tools/ref_engine_harness.py --layers 4 --tokens 8 --steps 16 --ratio 0.9 --dir gpurun_out/ref_store
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "moe-infinity_b200"))


def load_reference_engine():
    so = os.path.join(ROOT, "oracle", "_ref", "prefetch_op.so")
    if not os.path.exists(so):
        raise SystemExit(f"{so} missing: make -C oracle/ref_build -f Makefile.engine (needs /root/reference)")
    import torch  # noqa: F401  (libtorch first)
    spec = importlib.util.spec_from_file_location("prefetch_op", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--experts", type=int, default=8)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--inter", type=int, default=14336)
    ap.add_argument("--top-k", type=int, default=2)
    ap.add_argument("--tokens", type=int, default=8)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ratio", type=float, default=0.9, help="device_memory_ratio handed to prefetch_handle")
    ap.add_argument("--dir", default=os.path.join(ROOT, "gpurun_out", "ref_store"))
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--compare", type=int, default=1, help="also run this repo's engine and compare hidden states")
    args = ap.parse_args()

    import torch
    import torch.nn.functional as F
    P = load_reference_engine()
    L_, E, H, I, k, T = args.layers, args.experts, args.hidden, args.inter, args.top_k, args.tokens
    dt = torch.bfloat16
    os.makedirs(args.dir, exist_ok=True)

    # ---- 1. engine + tensor store (one handle per process)
    h = P.prefetch_handle(args.dir, args.ratio)
    g = torch.Generator().manual_seed(0)
    ids, next_id = {}, 0
    dense = []
    for i in range(2):                                           # two tiny dense stages (see gotchas)
        t = torch.randn(64, 64, generator=g).to(dt)
        h.offload(t, next_id)
        dense.append(next_id)
        next_id += 1
    experts = {}
    for l in range(L_):
        for e in range(E):
            ws = [(torch.randn(I, H, generator=g) * 0.02).to(dt), (torch.randn(H, I, generator=g) * 0.02).to(dt),
                  (torch.randn(I, H, generator=g) * 0.02).to(dt)]                  # w1, w2, w3 (expert_module.cpp:139-145)
            experts[(l, e)] = ws
            ids[(l, e)] = list(range(next_id, next_id + 3))
            for t, tid in zip(ws, ids[(l, e)]):
                h.offload(t, tid)                                                   # model_offload.py:894-899
            next_id += 3
    placeholders = {}
    for tid in range(next_id):
        placeholders[tid] = torch.zeros(1, dtype=dt)                               # model_offload.py:755,764
        h.register(placeholders[tid], tid)

    # ---- 2. dispatcher first, then the topology (real order), then expert registration
    d = P.expert_dispatcher(E, L_, 0, 4, args.threads)                             # dtype 0 = bf16, type 4 = Mixtral
    topology = [(f"dense{i}", [[dense[i]]]) for i in range(2)]
    topology += [(f"layer{l}", [ids[(l, e)] for e in range(E)]) for l in range(L_)]
    h.set_topology(topology)                                                       # model_offload.py:767-768
    for (l, e), t in ids.items():
        d.register_expert(l, e, t)                                                 # :851-853

    gates = [(torch.randn(E, H, generator=g) * 0.05).to(dt).cuda() for _ in range(L_)]
    xs = [torch.randn(T, H, generator=g).to(dt).cuda() for _ in range(L_)]

    def ref_layer(l, x):
        """mixtral.py:44-101 with dispatch_local (expert_executor.py:32-58) on the reference engine"""
        logits = F.linear(x, gates[l])
        rw = F.softmax(logits, dim=1, dtype=torch.float)
        rw, sel = torch.topk(rw, k, dim=-1)
        rw = (rw / rw.sum(dim=-1, keepdim=True)).to(x.dtype)
        mask = F.one_hot(sel, num_classes=E)
        wmask = (rw[..., None] * mask).permute(0, 2, 1).sum(dim=-1)
        mask = mask.permute(0, 2, 1).sum(dim=-1).to(torch.bool)
        count = mask.sum(dim=0).cpu().numpy().flatten()                            # the reference's per-layer host sync
        active = [e for e in range(E) if count[e] > 0]
        d.set_inputs(x, mask)
        d.set_expected_queue(len(active))
        for e in active:
            d.enqueue_expert(l, e, 0, False)
        out = torch.zeros_like(x)
        for y, _, e, _hit in d.wait_expert():
            idx = mask[:, e]
            out[idx] += y.to(x.device) * wmask[idx, e][:, None]
        return out, logits

    def run_ref():
        return [ref_layer(l, xs[l]) for l in range(L_)]

    for _ in range(args.warmup):
        run_ref()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outs = run_ref()
    torch.cuda.synchronize()
    ref_ms = (time.perf_counter() - t0) / args.steps * 1e3
    res = {"impl": "reference native engine (prefetch_op.so)", "layers": L_, "tokens": T, "ratio": args.ratio,
           "ms_per_step": ref_ms, "ms_per_layer": ref_ms / L_, "tokens_per_s_32_layers": T / (ref_ms / L_ * 32 / 1e3)}

    # ---- 3. this repository's engine on the same weights and the reference's own router logits
    if args.compare:
        from moe_infinity_b200 import MoEEngine
        eng = MoEEngine(num_layers=L_, num_experts=E, hidden=H, inter=I, top_k=k, dtype=dt, max_tokens=max(T, 16),
                        num_slots=L_ * E)
        for (l, e), ws in experts.items():
            eng.load_expert(l, e, ws)
        worst = 0.0
        for l in range(L_):
            ref_out, logits = outs[l]
            ours = eng.forward(l, xs[l], router_logits=logits)
            torch.cuda.synchronize()
            rms = ref_out.float().pow(2).mean().sqrt().item()
            worst = max(worst, (ours.float() - ref_out.float()).abs().max().item() / max(rms, 1e-9))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            for l in range(L_):
                eng.forward(l, xs[l])
        ev1.record()
        torch.cuda.synchronize()
        ours_ms = ev0.elapsed_time(ev1) / args.steps
        res.update({"ours_ms_per_step": ours_ms, "speedup": ref_ms / ours_ms, "max_abs_diff_over_rms": worst})
    print(json.dumps(res), flush=True)
    os._exit(0)            # never run the reference's destructors (see gotchas)


if __name__ == "__main__":
    main()

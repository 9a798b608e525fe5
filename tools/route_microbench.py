"""Warm-cache timing of the routing / combine kernels in isolation (CUDA events around a CUDA graph of 200 calls)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "moe-infinity_b200"))
import torch
from moe_infinity_b200 import MoEEngine, _lib as L


def t(fn, n=200):
    """per-call time in us of `fn` replayed n times inside ONE CUDA graph (what a kernel costs inside the step graph:
    its run time + the dependent-launch gap; no Python / driver launch cost)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3


def run(name, **kw):
    T = kw.pop("T")
    eng = MoEEngine(num_layers=1, max_tokens=max(T, 16), num_slots=kw["num_experts"], **kw)
    for e in range(kw["num_experts"]):
        eng.load_expert(0, e).normal_(0, 0.02)
    if kw.get("shared_inter"):
        eng._ck(eng.lib.b2m_register_shared(eng._h, 0, None, 0))
    gdt = torch.bfloat16 if kw.get("router", 0) == 0 else torch.float32
    eng.set_gate(0, torch.randn(kw["num_experts"], kw["hidden"], device="cuda").to(gdt) * 0.05)
    x = torch.randn(T, kw["hidden"], device="cuda").to(torch.bfloat16)
    out = torch.empty_like(x)
    lg = torch.randn(T, kw["num_experts"], device="cuda").to(gdt)
    eng.forward(0, x, out=out)
    print(name, "route(fused gate)  us:", round(t(lambda: eng.route(0, x)), 2))
    print(name, "route(logits given) us:", round(t(lambda: eng.route(0, x, router_logits=lg)), 2))
    eng.route(0, x); eng.run_experts(0, T)
    print(name, "combine            us:", round(t(lambda: eng.combine(0, x, out=out)), 2))
    print(name, "experts (K3+K4)    us:", round(t(lambda: eng.run_experts(0, T), 50), 2))
    print(name, "forward            us:", round(t(lambda: eng.forward(0, x, out=out), 50), 2))


run("mixtral T=8 ", T=8, num_experts=8, hidden=4096, inter=14336, top_k=2)
run("deepseek T=16", T=16, num_experts=64, hidden=2048, inter=1408, top_k=6, expert_type=L.EXPERT_DEEPSEEK,
    router=L.ROUTER_DEEPSEEK_GREEDY, shared_inter=2816)
# what do the shared experts (side stream + join) cost a DeepSeek decode layer?
run("deepseek T=16 (no shared experts)", T=16, num_experts=64, hidden=2048, inter=1408, top_k=6,
    expert_type=L.EXPERT_DEEPSEEK, router=L.ROUTER_DEEPSEEK_GREEDY)

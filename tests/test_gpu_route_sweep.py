"""GPU: randomized sweep of the routing kernels (K1/K2) over expert counts, top-k, token counts and logit dtypes --
both the small-T (two-kernel) and the large-T (multi-CTA scan) paths -- against the oracle's routing functions;
plus ragged model dimensions (H, I not multiples of the 64/128 tile sizes) and the graph-captured DecodeSession."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import moe_oracle as O  # noqa: E402
from test_gpu_parity import check_indices, check_permutation, check_weights, hidden_close  # noqa: E402


def _engine(E, k, H, router, dtype=torch.bfloat16, max_tokens=1024, **kw):
    from moe_infinity_b200 import MoEEngine, _lib as L
    et = L.EXPERT_MIXTRAL if router == L.ROUTER_MIXTRAL else L.EXPERT_DEEPSEEK
    return MoEEngine(num_layers=1, num_experts=E, hidden=H, inter=128, top_k=k, dtype=dtype, expert_type=et,
                     router=router, max_tokens=max_tokens, num_slots=1, **kw)


@pytest.mark.parametrize("E,k", [(2, 1), (3, 2), (8, 2), (8, 8), (16, 4), (33, 5), (64, 6), (128, 8), (256, 3)])
@pytest.mark.parametrize("T", [1, 31, 33, 256, 257, 700])
def test_mixtral_routing_sweep(E, k, T, lib_built):
    from moe_infinity_b200 import _lib as L
    H = 64
    eng = _engine(E, k, H, L.ROUTER_MIXTRAL)
    g = torch.Generator().manual_seed(E * 1000 + k * 10 + T)
    x = torch.randn(T, H, generator=g).to(torch.bfloat16)
    logits = (torch.randn(T, E, generator=g) * 2).to(torch.bfloat16)
    r = O.mixtral_route(logits, k, torch.bfloat16)
    eng.route(0, x.cuda(), router_logits=logits.cuda())
    torch.cuda.synchronize()
    check_indices(eng.ws("topk_idx", T), r.topk_idx, O.tied_tokens(r.scores, k))
    ok = ~O.tied_tokens(r.scores, k)
    check_weights(eng.ws("topk_w", T)[ok.cuda()], r.topk_weight[ok], torch.bfloat16)
    check_permutation(eng, x.cuda(), T)
    s = eng.ws("scores", T).cpu()
    assert torch.allclose(s, r.scores, rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize("E,k,n_group,topk_group", [(16, 4, 4, 2), (64, 6, 8, 3), (32, 8, 1, 1)])
@pytest.mark.parametrize("T", [5, 300])
@pytest.mark.parametrize("norm", [False, True])
def test_deepseek_routing_sweep(E, k, n_group, topk_group, T, norm, lib_built):
    from moe_infinity_b200 import _lib as L
    H = 64
    router = L.ROUTER_DEEPSEEK_GROUP if n_group > 1 else L.ROUTER_DEEPSEEK_GREEDY
    eng = _engine(E, k, H, router, n_group=n_group, topk_group=topk_group, norm_topk_prob=norm,
                  routed_scaling_factor=1.0 if norm else 3.0)
    g = torch.Generator().manual_seed(E + k + T)
    x = torch.randn(T, H, generator=g).to(torch.bfloat16)
    scores = torch.softmax(torch.randn(T, E, generator=g) * 2, dim=-1)
    r = O.deepseek_route(scores, k, topk_method="group_limited_greedy" if n_group > 1 else "greedy", n_group=n_group,
                         topk_group=topk_group, norm_topk_prob=norm, routed_scaling_factor=1.0 if norm else 3.0)
    eng.route(0, x.cuda(), scores=scores.cuda())
    torch.cuda.synchronize()
    gi, gw = eng.ws("topk_idx", T).cpu().long(), eng.ws("topk_w", T).cpu()
    # ties in the *selection* scores (zeros of masked groups can tie): compare where the selected values are distinct
    sel = torch.gather(scores, 1, r.topk_idx)
    distinct = (sel.sort(-1).values.diff(dim=-1) != 0).all(-1) & (sel > 0).all(-1) & ~O.tied_tokens(r.scores, k)
    assert distinct.float().mean() > 0.9
    assert torch.equal(gi[distinct].sort(-1).values, r.topk_idx[distinct].sort(-1).values)
    wm_gpu = torch.zeros(T, E).scatter_(1, gi, gw)
    wm_ref = torch.zeros(T, E).scatter_(1, r.topk_idx, r.topk_weight.float())
    assert torch.all((wm_gpu[distinct] - wm_ref[distinct]).abs() <= 1e-6 * wm_ref[distinct].abs() + 1e-12)
    check_permutation(eng, x.cuda(), T)


def test_ragged_model_dims(lib_built):
    """H=200, I=328: K tails are zero-filled by TMA, partial weight-row tiles are masked in the epilogue."""
    from test_gpu_parity import make_engine
    T, E, k, H, I = 45, 4, 2, 200, 328
    dt = torch.bfloat16
    experts = O.make_experts(E, H, I, dt, seed=9, std=0.05)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(1, T, H, generator=g).to(dt)
    gate = (torch.randn(E, H, generator=g) * 0.3).to(dt)
    ref, logits, r = O.mixtral_block(x, gate, experts, k)
    y32 = O.combine_fp32(x, experts, r.topk_idx, r.topk_weight, O.MIXTRAL_MOE_DENSE_ACT_DENSE)
    c = dict(B=1, S=T, E=E, H=H, I=I, k=k, dtype=dt, experts=experts, gate=gate)
    for impl in (0, 1):
        eng = make_engine(c, "mixtral", gemm_impl=impl)
        out = eng.forward(0, x.cuda(), router_logits=logits.cuda())
        torch.cuda.synchronize()
        rows = (eng.ws("topk_w", T).cpu() == r.topk_weight.float()).all(-1) & ~O.tied_tokens(r.scores, k)
        hidden_close(out.reshape(T, -1)[rows.cuda()], ref.reshape(T, -1)[rows], y32.reshape(T, -1)[rows], dt, f"ragged impl={impl}")


def test_decode_session_graph_matches_eager(lib_built):
    from moe_infinity_b200 import DecodeSession, MoEEngine
    L_, E, k, H, I, T = 3, 8, 2, 256, 384, 8
    dt = torch.bfloat16
    eng = MoEEngine(num_layers=L_, num_experts=E, hidden=H, inter=I, top_k=k, dtype=dt, max_tokens=16, num_slots=L_ * E)
    torch.manual_seed(0)
    for l in range(L_):
        for e in range(E):
            eng.load_expert(l, e).normal_(0, 0.05)
        eng.set_gate(l, torch.randn(E, H) * 0.3)
    sess = DecodeSession(eng, T).capture()
    for it in range(3):
        sess.x_host.copy_(torch.randn(L_, T, H).to(dt))
        out = sess.step().clone()
        for l in range(L_):
            want = eng.forward(l, sess.x_host[l].cuda())
            torch.cuda.synchronize()
            assert torch.equal(out[l].cuda(), want), (it, l)
    assert eng.stats()["host_syncs"] == 0


@pytest.mark.parametrize("router", ["mixtral", "deepseek"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("E,k,H,T", [(8, 2, 4096, 300), (12, 3, 200, 257), (64, 6, 2048, 700), (33, 5, 72, 1000),
                                     (256, 8, 264, 513), (8, 2, 1000, 19), (64, 6, 200, 16)])
def test_fused_gate_sweep(router, dtype, E, k, H, T, lib_built):
    """K0 computed in-kernel (no router input): the prefill gate GEMM (T > 256: cp.async-staged gate chunks, expert groups
    split over blockIdx.y, H and E tails) and the decode gate.  Logits vs an fp64 product: Mixtral rounds to the model
    dtype (mixtral.py:46) -> within 1 ulp; DeepSeek keeps fp32 (modeling_deepseek.py:467-471) -> rel 2e-5.
    Expert indices must equal the oracle's routing of the GPU's own logits exactly (ties excluded)."""
    from moe_infinity_b200 import _lib as L
    mix = router == "mixtral"
    eng = _engine(E, k, H, L.ROUTER_MIXTRAL if mix else L.ROUTER_DEEPSEEK_GREEDY, dtype=dtype, max_tokens=1024)
    g = torch.Generator().manual_seed(E * 7 + H + T)
    x = torch.randn(T, H, generator=g).to(dtype)
    w = (torch.randn(E, H, generator=g) * (2.0 / H ** 0.5)).to(dtype if mix else torch.float32)
    eng.set_gate(0, w)
    eng.route(0, x.cuda())
    torch.cuda.synchronize()
    lg = eng.ws("logits", T).cpu()
    exact = x.double() @ w.double().t()
    if mix:
        eps = torch.finfo(dtype).eps
        assert lg.dtype == dtype
        assert torch.all((lg.double() - exact).abs() <= eps * exact.abs() + 1e-4), "gate logits beyond 1 ulp"
        r = O.mixtral_route(lg, k, dtype)
        check_indices(eng.ws("topk_idx", T), r.topk_idx, O.tied_tokens(r.scores, k))
    else:
        assert lg.dtype == torch.float32
        assert torch.allclose(lg.double(), exact, rtol=2e-5, atol=2e-5)
        scores = torch.softmax(lg, dim=-1, dtype=torch.float32)
        r = O.deepseek_route(scores, k)
        near = O.tied_tokens(r.scores, k)
        gi = eng.ws("topk_idx", T).cpu().long()
        # softmax on the GPU differs from CPU torch by <= 2 ulp(fp32): exclude tokens whose k-th/(k+1)-th scores are that close
        srt = scores.sort(-1, descending=True).values
        if E > k:
            near |= (srt[:, k - 1] - srt[:, k]).abs() <= 4e-7 * srt[:, k - 1]
        assert near.float().mean() < 0.05
        assert torch.equal(gi[~near].sort(-1).values, r.topk_idx[~near].sort(-1).values)
    check_permutation(eng, x.cuda(), T)


@pytest.mark.parametrize("E,k", [(8, 2), (64, 6), (256, 8), (33, 5)])
@pytest.mark.parametrize("T", [1, 32, 33, 64, 100, 200, 256])
def test_decode_routing_all_resident_sweep(E, k, T, lib_built):
    """All experts resident and T <= 256: the decode routing path of the fused call -- the last gate/top-k CTA publishes
    counts, offsets AND the row maps (count/rank spread over its warps, 1..8 token chunks), the permute kernel only copies
    rows.  (The sweeps above run with one HBM slot, i.e. in offload mode, where every permute CTA ranks for itself.)"""
    from moe_infinity_b200 import MoEEngine, _lib as L
    H = 64
    eng = MoEEngine(num_layers=1, num_experts=E, hidden=H, inter=128, top_k=k, dtype=torch.bfloat16,
                    expert_type=L.EXPERT_MIXTRAL, router=L.ROUTER_MIXTRAL, max_tokens=256, num_slots=E)
    for e in range(E):
        eng.load_expert(0, e).zero_()
    assert eng.stats()["slots"] == E
    g = torch.Generator().manual_seed(E * 131 + k * 17 + T)
    x = torch.randn(T, H, generator=g).to(torch.bfloat16)
    logits = (torch.randn(T, E, generator=g) * 2).to(torch.bfloat16)
    r = O.mixtral_route(logits, k, torch.bfloat16)
    eng.route(0, x.cuda(), router_logits=logits.cuda())
    torch.cuda.synchronize()
    tied = O.tied_tokens(r.scores, k)
    check_indices(eng.ws("topk_idx", T), r.topk_idx, tied)
    check_permutation(eng, x.cuda(), T)
    if not bool(tied.any()):
        counts = torch.bincount(r.topk_idx.flatten(), minlength=E)
        assert eng.ws("counts", T).cpu().tolist() == counts.tolist()
        assert eng.ws("offsets", T).cpu().tolist() == [0] + torch.cumsum(counts, 0).tolist()
    assert eng.stats()["host_syncs"] == 0

"""GPU: edge cases of the dispatch path -- empty input, one token, every token to the same experts, k == E,
many experts (E=128, k=8), workspace-capacity boundary, argument validation through the C-ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import moe_oracle as O  # noqa: E402
from test_gpu_parity import check_indices, check_permutation, hidden_close, make_engine  # noqa: E402


def _case(T, H, I, E, k, dtype=torch.bfloat16, seed=0, gate_scale=0.3):
    experts = O.make_experts(E, H, I, dtype, seed=seed, std=0.05)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(1, T, H, generator=g).to(dtype)
    gate = (torch.randn(E, H, generator=g) * gate_scale).to(dtype)
    return dict(B=1, S=T, E=E, H=H, I=I, k=k, dtype=dtype, experts=experts, gate=gate), x


def test_empty_input(lib_built):
    c, _ = _case(4, 128, 256, 8, 2)
    eng = make_engine(c, "mixtral")
    x = torch.zeros(0, 128, dtype=torch.bfloat16, device="cuda")
    out = eng.forward(0, x)
    torch.cuda.synchronize()
    assert out.shape == (0, 128)
    assert eng.ws("offsets", 0).cpu().tolist() == [0] * 9


def test_all_tokens_same_experts_multi_ntile(lib_built):
    """Constant logits -> every token ties -> experts (0,1); 300 tokens per expert spans several N tiles."""
    T = 300
    c, x = _case(T, 128, 256, 8, 2)
    logits = torch.zeros(T, 8, dtype=torch.bfloat16)
    ref, _, r = O.mixtral_block(x, None, c["experts"], 2, router_logits=logits)
    eng = make_engine(c, "mixtral")
    out = eng.forward(0, x.cuda(), router_logits=logits.cuda())
    torch.cuda.synchronize()
    assert eng.ws("counts", T).cpu().tolist() == [T, T, 0, 0, 0, 0, 0, 0]
    assert torch.equal(eng.ws("topk_idx", T).cpu().long(), r.topk_idx)      # ties -> lowest index, like the oracle
    hidden_close(out, ref, None, c["dtype"], "same experts")


def test_topk_equals_num_experts(lib_built):
    T, E = 20, 4
    c, x = _case(T, 128, 128, E, E, dtype=torch.float16)
    ref, logits, r = O.mixtral_block(x, c["gate"], c["experts"], E)
    eng = make_engine(c, "mixtral")
    out = eng.forward(0, x.cuda(), router_logits=logits.cuda())
    torch.cuda.synchronize()
    assert eng.ws("counts", T).cpu().tolist() == [T] * E
    rows = (eng.ws("topk_w", T).cpu() == r.topk_weight.float()).all(-1)
    # combine order is ascending expert id on both sides; k=4 chained roundings
    hidden_close(out.reshape(T, -1)[rows.cuda()], ref.reshape(T, -1)[rows], None, c["dtype"], "k==E")


def test_many_experts_deepseek_like(lib_built):
    """E=128, k=8 (the kernel maximum), greedy DeepSeek routing with fp32 weights, no shared expert."""
    T, E, k, H, I = 70, 128, 8, 128, 128
    dt = torch.bfloat16
    experts = O.make_experts(E, H, I, dt, seed=5, expert_type=O.DEEPSEEK_MOE_DENSE_ACT_DENSE, std=0.05)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, T, H, generator=g).to(dt)
    gate = torch.randn(E, H, generator=g) * 0.2
    ref, r = O.deepseek_block(x, gate, experts, k, None, routed_scaling_factor=2.5)
    c = dict(B=1, S=T, E=E, H=H, I=I, k=k, dtype=dt, experts=experts, gate=gate, topk_method="greedy", n_shared=None,
             n_group=1, topk_group=1, norm_topk_prob=False, routed_scaling_factor=2.5, shared=None)
    eng = make_engine(c, "deepseek")
    out = eng.forward(0, x.cuda(), scores=r.scores.cuda())
    torch.cuda.synchronize()
    check_indices(eng.ws("topk_idx", T), r.topk_idx, O.tied_tokens(r.scores, k), sort_rows=True)
    check_permutation(eng, x.cuda().reshape(T, -1), T)
    hidden_close(out, ref, None, dt, "E=128 k=8")


def test_workspace_capacity_boundary_and_errors(lib_built):
    from moe_infinity_b200 import B2MError
    c, x = _case(32, 128, 256, 8, 2)
    eng = make_engine(c, "mixtral", max_tokens=32)
    ref, logits, _ = O.mixtral_block(x, c["gate"], c["experts"], 2)
    out = eng.forward(0, x.cuda(), router_logits=logits.cuda())          # exactly max_tokens
    hidden_close(out, ref, None, c["dtype"], "T == max_tokens")
    with pytest.raises(B2MError):
        eng.forward(0, torch.zeros(33, 128, dtype=torch.bfloat16, device="cuda"))   # beyond capacity
    with pytest.raises(B2MError):
        eng.forward(3, x.cuda())                                          # layer out of range
    with pytest.raises(ValueError):
        eng.forward(0, x.cuda().float())                                  # wrong dtype
    with pytest.raises(B2MError):
        eng.run_experts(0, 5)                                             # no matching routing call
    # the context is still usable after errors
    out2 = eng.forward(0, x.cuda(), router_logits=logits.cuda())
    assert torch.equal(out, out2)


def test_unsupported_types_fail_loudly(lib_built):
    from moe_infinity_b200 import MoEEngine, B2MError, _lib as L
    with pytest.raises(B2MError):
        MoEEngine(num_layers=1, num_experts=8, hidden=128, inter=256, top_k=2, expert_type=6)   # expert_module.h: 0..5
    with pytest.raises(ValueError):
        MoEEngine(num_layers=1, num_experts=8, hidden=128, inter=256, top_k=2, dtype=torch.float8_e4m3fn)   # dtype int 3: no Float8 matmul in the reference either
    eng = MoEEngine(num_layers=1, num_experts=8, hidden=128, inter=256, top_k=2, dtype=torch.float32)       # dtype int 1: CUDA-core fp32 path
    with pytest.raises(B2MError):       # ... which takes router logits / scores / masks, not the fused 16-bit gate kernel
        eng.forward(0, torch.zeros(4, 128, device="cuda"))
    eng.close()
    with pytest.raises(B2MError):
        MoEEngine(num_layers=1, num_experts=8, hidden=128, inter=256, top_k=9)


def test_gelu_gated_switch_expert_type(lib_built):
    """expert_type 1 (SwitchTransformersDenseGatedActDense: gelu(x wi_0^T) * (x wi_1^T) wo^T, expert_module.cpp:54-59)."""
    from moe_infinity_b200 import MoEEngine, _lib as L
    T, E, H, I = 24, 4, 128, 256
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(3)
    experts = [[(torch.randn(s, generator=g) * 0.05).to(dt) for s in ((I, H), (I, H), (H, I))] for _ in range(E)]
    x = torch.randn(T, H, generator=g).to(dt)
    eng = MoEEngine(num_layers=1, num_experts=E, hidden=H, inter=I, top_k=1, dtype=dt, expert_type=L.EXPERT_SWITCH_GATED,
                    router=L.ROUTER_SWITCH_TOP1, expert_capacity=64, max_tokens=32)
    for e in range(E):
        eng.load_expert(0, e, experts[e])
    mask = torch.zeros(T, E, dtype=torch.uint8)
    mask[torch.arange(T), torch.arange(T) % E] = 1
    eng.route_from_mask(0, x.cuda(), mask.cuda())
    eng.run_experts(0, T)
    rows, offs = eng.expert_outputs(T)
    for e in range(E):
        want = O.expert_ffn(x[mask[:, e].bool()], experts[e], O.SWITCH_DENSE_GATED_ACT_DENSE)
        hidden_close(rows[offs[e]:offs[e + 1]], want, None, dt, f"gated expert {e}")

"""Generate tests/golden/*.pt from the LITERAL reference block files (dev container only).

Run:  python tests/golden/make_golden.py
Each fixture stores the seeded inputs' identity (seed + weight checksum), the router
logits/scores, and the outputs of the literal reference block
(/root/reference/moe_infinity/models/{mixtral,deepseek}.py + MoEGate) executed on CPU with a
pure-torch `expert_executor` stand-in that calls the block's own HF expert modules in
ascending expert order (the authors' commented loop, mixtral.py:103-113).
Weights are not stored (regenerated from the seed; a checksum guards generator drift).
Switch fixtures: the literal SyncSwitchTransformersSparseMLP (switch_transformers.py:41-113) runs on a 4.x-order router
shim (tests/shims/ref_loader.py shim 5: HF 5.5's router returns a different tuple), with HF's own
SwitchTransformersDenseActDense experts; generation asserts literal == oracle bit for bit.
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))

from oracle import moe_oracle as O  # noqa: E402

MIXTRAL_CASES = {
    # name: (H, I, E, k, B, S, dtype, seed)
    "mixtral_mini_bf16": (128, 256, 8, 2, 2, 12, torch.bfloat16, 11),
    "mixtral_mini_f16": (128, 256, 8, 2, 2, 12, torch.float16, 12),
    "mixtral_ragged_bf16": (192, 320, 8, 2, 1, 37, torch.bfloat16, 13),
    "mixtral_onetoken_bf16": (128, 256, 8, 2, 1, 1, torch.bfloat16, 14),
}
DEEPSEEK_CASES = {
    # name: (H, I, E, k, n_shared, B, S, dtype, seed, topk_method, n_group, topk_group, norm, scale)
    "deepseek_mini_bf16": (128, 128, 16, 4, 2, 2, 10, torch.bfloat16, 21, "greedy", 1, 1, False, 1.0),
    "deepseek_group_bf16": (128, 128, 16, 4, 2, 1, 19, torch.bfloat16, 22, "group_limited_greedy", 4, 2, False, 16.0),
    "deepseek_norm_f16": (128, 64, 16, 6, None, 1, 9, torch.float16, 23, "greedy", 1, 1, True, 1.0),
}
SWITCH_CASES = {
    # name: (D, F, E, capacity, B, S, dtype, seed)
    "switch_mini_f32": (64, 256, 8, 6, 1, 32, torch.float32, 31),
    "switch_mini_bf16": (128, 256, 8, 64, 2, 16, torch.bfloat16, 32),
}


NLLB_CASES = {
    # name: (D, F, E, capacity, B, S, dtype, seed)      literal block only (HF's NllbMoeTop2Router does the routing)
    "nllb_mini_bf16": (128, 256, 8, 64, 2, 12, torch.bfloat16, 41),
    "nllb_capacity_f16": (128, 128, 8, 5, 1, 24, torch.float16, 42),     # capacity 5 of 24 tokens: HF drops the overflow
}


def checksum(tensors) -> float:
    s = 0.0
    for t in tensors:
        s += float(t.double().abs().sum())
    return s


def gen_hidden(B, S, H, dtype, seed):
    g = torch.Generator().manual_seed(seed + 1000)
    return torch.randn(B, S, H, generator=g).to(dtype)


def gen_gate(E, H, dtype, seed, std=0.5):
    g = torch.Generator().manual_seed(seed + 2000)
    return (torch.randn(E, H, generator=g) * std / (H ** 0.5)).to(dtype)


class _StandInExecutor:
    """dispatch_local stand-in: the literal block's own experts, ascending expert id."""

    def __init__(self, block, experts_attr="experts"):
        self.block = block
        self.experts_attr = experts_attr

    def dispatch_local(self, hidden_states, router_mask, layer_id):
        E = router_mask.shape[-1]
        experts = getattr(self.block, self.experts_attr)
        x = hidden_states.reshape(-1, hidden_states.shape[-1])
        m = router_mask.reshape(-1, E)
        out = []
        for e in range(E):
            idx = m[:, e].to(torch.bool)
            if bool(idx.any()):
                out.append((experts[e](x[idx]), layer_id, e, 1))
        return out


def run_literal_mixtral(ns, H, I, E, k, hidden, gate_w, experts):
    cfg = types.SimpleNamespace(hidden_size=H, intermediate_size=I, num_local_experts=E,
                                num_experts_per_tok=k, hidden_act="silu")
    blk = ns.mixtral.SyncMixtralSparseMoeBlock(cfg).to(hidden.dtype)
    with torch.no_grad():
        blk.gate.weight.copy_(gate_w)
        for e in range(E):
            blk.experts[e].w1.weight.copy_(experts[e][0])
            blk.experts[e].w2.weight.copy_(experts[e][1])
            blk.experts[e].w3.weight.copy_(experts[e][2])
    blk.expert_executor = _StandInExecutor(blk)
    blk.layer_id = 0
    with torch.no_grad():
        out, logits = blk(hidden)
    return out, logits


def run_literal_deepseek(ns, H, I, E, k, n_shared, hidden, gate_w, experts, shared, method, n_group, topk_group,
                         norm, scale):
    cfg = types.SimpleNamespace(model_type="deepseek_v2", hidden_size=H, intermediate_size=I * 4,
                                moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k,
                                n_shared_experts=n_shared, routed_scaling_factor=scale, scoring_func="softmax",
                                aux_loss_alpha=0.0, seq_aux=False, topk_method=method, n_group=n_group,
                                topk_group=topk_group, norm_topk_prob=norm, hidden_act="silu",
                                pretraining_tp=1)
    blk = ns.deepseek.DeepseekMoEBlock(cfg)
    blk = blk.to(hidden.dtype)
    blk.eval()
    with torch.no_grad():
        blk.gate.weight.copy_(gate_w)
        for e in range(E):
            blk.experts[e].gate_proj.weight.copy_(experts[e][0])
            blk.experts[e].up_proj.weight.copy_(experts[e][1])
            blk.experts[e].down_proj.weight.copy_(experts[e][2])
        if n_shared is not None:
            blk.shared_experts.gate_proj.weight.copy_(shared[0])
            blk.shared_experts.up_proj.weight.copy_(shared[1])
            blk.shared_experts.down_proj.weight.copy_(shared[2])
    blk.expert_executor = _StandInExecutor(blk)
    blk.layer_id = 0
    with torch.no_grad():
        out = blk(hidden)
    return out


def run_literal_switch(ns, D, Fd, E, cap, hidden, gate_w, experts):
    from transformers import SwitchTransformersConfig
    cfg = SwitchTransformersConfig(d_model=D, d_ff=Fd, num_experts=E, expert_capacity=cap, router_bias=False,
                                   router_jitter_noise=0.0, router_dtype="float32", dropout_rate=0.0,
                                   dense_act_fn="relu", num_layers=1, num_sparse_encoder_layers=1)
    blk = ns.switch.SyncSwitchTransformersSparseMLP(cfg)
    blk = blk.to(hidden.dtype)
    blk.eval()
    with torch.no_grad():
        blk.router.classifier.weight.copy_(gate_w)
        for e in range(E):
            blk.experts[f"expert_{e}"].wi.weight.copy_(experts[e][0])
            blk.experts[f"expert_{e}"].wo.weight.copy_(experts[e][1])
    ex = _StandInExecutor(blk)
    ex.dispatch_local = lambda h, m, lid: [(blk.experts[f"expert_{e}"](h.reshape(-1, h.shape[-1])[m.reshape(-1, E)[:, e].bool()]),
                                            lid, e, 1) for e in range(E) if bool(m.reshape(-1, E)[:, e].any())]
    blk.expert_executor = ex
    blk.layer_id = 0
    # the block's last line moves its auxiliary outputs to "cuda:0" (switch_transformers.py:110-113): on CPU, stand in
    orig_to = torch.Tensor.to

    def _to(self, *a, **kw):
        if a and isinstance(a[0], str) and a[0].startswith("cuda") and not torch.cuda.is_available():
            return self
        return orig_to(self, *a, **kw)
    torch.Tensor.to = _to
    try:
        with torch.no_grad():
            out, (logits, expert_index) = blk(hidden)
    finally:
        torch.Tensor.to = orig_to
    return out, logits, expert_index


def nllb_config(D, Fd, E, cap):
    from transformers import NllbMoeConfig
    return NllbMoeConfig(d_model=D, encoder_ffn_dim=Fd, decoder_ffn_dim=Fd, num_experts=E, expert_capacity=cap,
                         router_bias=False, router_dtype="float32", router_jitter_noise=0.0, moe_token_dropout=0.0,
                         dropout=0.0, activation_dropout=0.0, activation_function="relu", second_expert_policy="all",
                         normalize_router_prob_before_dropping=False, batch_prioritized_routing=False,
                         moe_eval_capacity_token_fraction=-1.0)


def run_literal_nllb(ns, D, Fd, E, cap, hidden, gate_w, experts):
    """The literal SyncNllbMoeSparseMLP (nllb_moe.py:20-115) in eval mode; behind dispatch_local: the oracle's statement of
    the engine's NllbMoeDenseActDense module (expert_module.cpp:79-129, pinned on the compiled reference module)."""
    blk = ns.nllb.SyncNllbMoeSparseMLP(nllb_config(D, Fd, E, cap), Fd).to(hidden.dtype)
    blk.eval()
    with torch.no_grad():
        blk.router.classifier.weight.copy_(gate_w)
    ex = types.SimpleNamespace()
    ex.dispatch_local = lambda h, m, lid: O.dispatch_local(h, m, experts, O.NLLB_MOE_DENSE_ACT_DENSE, lid)
    blk.expert_executor, blk.layer_id = ex, 0
    orig_to = torch.Tensor.to

    def _to(self, *a, **kw):                      # the block's last lines move its auxiliary outputs to "cuda:0" (:111-114)
        if a and isinstance(a[0], str) and a[0].startswith("cuda") and not torch.cuda.is_available():
            return self
        return orig_to(self, *a, **kw)
    torch.Tensor.to = _to
    try:
        with torch.no_grad():
            out, (router_probs, top1) = blk(hidden)
    finally:
        torch.Tensor.to = orig_to
    return out, router_probs, top1


def build_nllb(name):
    D, Fd, E, cap, B, S, dtype, seed = NLLB_CASES[name]
    experts = O.make_experts(E, D, Fd, dtype, seed, O.NLLB_MOE_DENSE_ACT_DENSE, std=0.05)
    return dict(H=D, I=Fd, E=E, capacity=cap, B=B, S=S, dtype=dtype, seed=seed, experts=experts,
                hidden=gen_hidden(B, S, D, dtype, seed),
                gate=gen_gate(E, D, dtype, seed, std=2.0).float())      # fp32 values representable in the model dtype


def make_nllb(ns):
    for name in NLLB_CASES:
        c = build_nllb(name)
        out, probs, top1 = run_literal_nllb(ns, c["H"], c["I"], c["E"], c["capacity"], c["hidden"], c["gate"], c["experts"])
        torch.save(dict(kind="nllb", source="literal",
                        cfg={k: c[k] for k in ("H", "I", "E", "capacity", "B", "S", "seed")}, dtype=str(c["dtype"]),
                        weight_checksum=checksum([w for e in c["experts"] for w in e]), hidden=c["hidden"],
                        gate=c["gate"], router_probs=probs, top1=top1, out=out), os.path.join(HERE, name + ".pt"))
        print(name, "literal ok; experts per token:", sorted(set((probs != 0).sum(-1).flatten().tolist())))


def build_mixtral(name):
    H, I, E, k, B, S, dtype, seed = MIXTRAL_CASES[name]
    experts = O.make_experts(E, H, I, dtype, seed, O.MIXTRAL_MOE_DENSE_ACT_DENSE, std=0.05)
    return dict(H=H, I=I, E=E, k=k, B=B, S=S, dtype=dtype, seed=seed, experts=experts,
                hidden=gen_hidden(B, S, H, dtype, seed), gate=gen_gate(E, H, dtype, seed, std=2.0))


def build_deepseek(name):
    H, I, E, k, ns_, B, S, dtype, seed, method, n_group, topk_group, norm, scale = DEEPSEEK_CASES[name]
    experts = O.make_experts(E, H, I, dtype, seed, O.DEEPSEEK_MOE_DENSE_ACT_DENSE, std=0.05)
    shared = None
    if ns_ is not None:
        shared = O.make_experts(1, H, I * ns_, dtype, seed + 500, O.DEEPSEEK_MOE_DENSE_ACT_DENSE, std=0.05)[0]
    return dict(H=H, I=I, E=E, k=k, n_shared=ns_, B=B, S=S, dtype=dtype, seed=seed, experts=experts,
                shared=shared, hidden=gen_hidden(B, S, H, dtype, seed), gate=gen_gate(E, H, dtype, seed, std=2.0),
                topk_method=method, n_group=n_group, topk_group=topk_group, norm_topk_prob=norm,
                routed_scaling_factor=scale)


def build_switch(name):
    D, Fd, E, cap, B, S, dtype, seed = SWITCH_CASES[name]
    experts = O.make_experts(E, D, Fd, dtype, seed, O.SWITCH_DENSE_ACT_DENSE, std=0.05)
    return dict(H=D, I=Fd, E=E, capacity=cap, B=B, S=S, dtype=dtype, seed=seed, experts=experts,
                hidden=gen_hidden(B, S, D, dtype, seed),
                # the classifier is a parameter of a `dtype` model that the router up-casts to fp32 per call (HF
                # _cast_classifier): its fp32 values are exactly representable in the model dtype
                gate=gen_gate(E, D, dtype, seed, std=2.0).float())


def main():
    import ref_loader
    ns = ref_loader.load() if ref_loader.available() else None
    if sys.argv[1:] == ["nllb"]:                  # only the NLLB fixtures (the others are left untouched)
        make_nllb(ns)
        return
    if ns is None:
        print("WARNING: reference tree absent; fixtures will come from the oracle only")
    for name in MIXTRAL_CASES:
        c = build_mixtral(name)
        o_out, o_logits, r = O.mixtral_block(c["hidden"], c["gate"], c["experts"], c["k"])
        src = "oracle"
        if ns is not None:
            l_out, l_logits = run_literal_mixtral(ns, c["H"], c["I"], c["E"], c["k"], c["hidden"], c["gate"],
                                                  c["experts"])
            tied = O.tied_tokens(r.scores, c["k"])
            assert torch.equal(l_logits, o_logits), name
            ok_rows = ~tied
            lo = l_out.reshape(-1, c["H"])
            oo = o_out.reshape(-1, c["H"])
            assert torch.equal(lo[ok_rows], oo[ok_rows]), f"{name}: literal != oracle"
            o_out, src = l_out, "literal"
        y32 = O.combine_fp32(c["hidden"], c["experts"], r.topk_idx, r.topk_weight, O.MIXTRAL_MOE_DENSE_ACT_DENSE)
        torch.save(dict(kind="mixtral", source=src, cfg={k: c[k] for k in ("H", "I", "E", "k", "B", "S", "seed")},
                        dtype=str(c["dtype"]), weight_checksum=checksum([w for e in c["experts"] for w in e]),
                        hidden=c["hidden"], gate=c["gate"], router_logits=o_logits, topk_idx=r.topk_idx,
                        topk_weight=r.topk_weight, tied=O.tied_tokens(r.scores, c["k"]), out=o_out, out_fp32=y32),
                   os.path.join(HERE, name + ".pt"))
        print(name, src, "ok")
    for name in DEEPSEEK_CASES:
        c = build_deepseek(name)
        kw = dict(topk_method=c["topk_method"], n_group=c["n_group"], topk_group=c["topk_group"],
                  norm_topk_prob=c["norm_topk_prob"], routed_scaling_factor=c["routed_scaling_factor"])
        o_out, r = O.deepseek_block(c["hidden"], c["gate"], c["experts"], c["k"], c["shared"], **kw)
        src = "oracle"
        if ns is not None:
            l_out = run_literal_deepseek(ns, c["H"], c["I"], c["E"], c["k"], c["n_shared"], c["hidden"], c["gate"],
                                         c["experts"], c["shared"], c["topk_method"], c["n_group"], c["topk_group"],
                                         c["norm_topk_prob"], c["routed_scaling_factor"])
            tied = O.tied_tokens(r.scores, c["k"])
            lo = l_out.reshape(-1, c["H"])
            oo = o_out.reshape(-1, c["H"])
            if c["norm_topk_prob"]:
                # torch.topk(sorted=False) returns the k experts in an unspecified order and the reference
                # sums the k weights in that order (modeling_deepseek.py:508-510): the denominator, hence the
                # weights, are order dependent in the last fp32 bit.  The oracle fixes descending-score order;
                # the literal run must agree to within one ulp of the output dtype.
                ulp = torch.finfo(c["dtype"]).eps
                assert torch.all((lo.float() - oo.float()).abs() <= ulp * oo.float().abs() + 1e-6), name
                src = "oracle (literal agrees within 1 ulp; topk order dependent)"
            else:
                assert torch.equal(lo[~tied], oo[~tied]), f"{name}: literal != oracle"
                o_out, src = l_out, "literal"
        y32 = O.combine_fp32(c["hidden"], c["experts"], r.topk_idx, r.topk_weight, O.DEEPSEEK_MOE_DENSE_ACT_DENSE,
                             c["shared"])
        torch.save(dict(kind="deepseek", source=src,
                        cfg={k: c[k] for k in ("H", "I", "E", "k", "n_shared", "B", "S", "seed", "topk_method",
                                               "n_group", "topk_group", "norm_topk_prob", "routed_scaling_factor")},
                        dtype=str(c["dtype"]), weight_checksum=checksum([w for e in c["experts"] for w in e]),
                        hidden=c["hidden"], gate=c["gate"], scores=r.scores, topk_idx=r.topk_idx,
                        topk_weight=r.topk_weight, tied=O.tied_tokens(r.scores, c["k"]), out=o_out, out_fp32=y32),
                   os.path.join(HERE, name + ".pt"))
        print(name, src, "ok")
    for name in SWITCH_CASES:
        c = build_switch(name)
        out, (logits, expert_index), router_mask = O.switch_block(c["hidden"], c["gate"], c["experts"], c["capacity"])
        src = "oracle"
        if ns is not None and hasattr(ns, "switch"):
            l_out, l_logits, l_index = run_literal_switch(ns, c["H"], c["I"], c["E"], c["capacity"], c["hidden"], c["gate"],
                                                          c["experts"])
            assert torch.equal(l_logits.float(), logits.float()), f"{name}: literal router logits != oracle"
            assert torch.equal(l_index, expert_index), f"{name}: literal expert index != oracle"
            assert torch.equal(l_out, out), f"{name}: literal != oracle"
            out, src = l_out, "literal"
        torch.save(dict(kind="switch", source=src,
                        cfg={k: c[k] for k in ("H", "I", "E", "capacity", "B", "S", "seed")}, dtype=str(c["dtype"]),
                        weight_checksum=checksum([w for e in c["experts"] for w in e]), hidden=c["hidden"],
                        gate=c["gate"], router_logits=logits, expert_index=expert_index, router_mask=router_mask,
                        out=out), os.path.join(HERE, name + ".pt"))
        print(name, src, "ok")


if __name__ == "__main__":
    main()

"""GPU: the reference-facing plugin surface -- drop-in blocks, `expert_dispatcher`/`prefetch_handle`/
`DistributedExpertExecutor` mirrors -- against the golden vectors of the literal reference blocks."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import moe_oracle as O  # noqa: E402
from test_gpu_parity import hidden_close, load_case, make_engine  # noqa: E402


def test_mixtral_block_fast_path_golden(lib_built):
    from moe_infinity_b200.blocks import SyncMixtralSparseMoeBlock
    from moe_infinity_b200.compat import DistributedExpertExecutor
    c, fx = load_case("mixtral_mini_bf16")
    eng = make_engine(c, "mixtral")
    cfg = types.SimpleNamespace(hidden_size=c["H"], intermediate_size=c["I"], num_local_experts=c["E"],
                                num_experts_per_tok=c["k"])
    blk = SyncMixtralSparseMoeBlock(cfg).to(torch.bfloat16).cuda()
    with torch.no_grad():
        blk.gate.weight.copy_(c["gate"])
    ex = DistributedExpertExecutor()
    ex.set_expert_dispatcher(eng)
    blk.expert_executor, blk.layer_id = ex, 0
    with torch.no_grad():
        out, logits = blk(c["hidden"].cuda())
    assert out.shape == c["hidden"].shape and logits.shape == (c["B"] * c["S"], c["E"])
    same = (logits.cpu() == fx["router_logits"]).all(dim=-1) & ~fx["tied"]    # GPU vs CPU gate GEMM may round differently
    assert same.float().mean() > 0.9
    T = c["B"] * c["S"]
    hidden_close(out.reshape(T, -1)[same.cuda()], fx["out"].reshape(T, -1)[same], None, c["dtype"], "block")


def test_reference_compat_objects_dispatch_local(lib_built):
    """prefetch_handle.offload -> expert_dispatcher.register_expert(ids) -> dispatch_local, then the reference's
    own Python combine loop (compat path of the block)."""
    from moe_infinity_b200 import _lib as L
    from moe_infinity_b200.blocks import SyncMixtralSparseMoeBlock
    from moe_infinity_b200.compat import DistributedExpertExecutor, expert_dispatcher, prefetch_handle
    c, fx = load_case("mixtral_mini_bf16")
    E = c["E"]
    h = prefetch_handle("/tmp/b2m_unused", 0.5)
    tid = 0
    ids = {}
    for e in range(E):
        ids[e] = []
        for w in c["experts"][e]:       # w1, w2, w3 in named_parameters order
            h.offload(w, tid)
            ids[e].append(tid)
            tid += 1
    d = expert_dispatcher(E, 1, L.DTYPE_BF16, L.EXPERT_MIXTRAL, 8, handle=h, top_k=c["k"], max_tokens=64, num_slots=E)
    for e in range(E):
        d.register_expert(0, e, ids[e])
    ex = DistributedExpertExecutor()
    ex.set_expert_dispatcher(d)
    r = O.mixtral_route(fx["router_logits"], c["k"], c["dtype"])
    x = c["hidden"].reshape(-1, c["H"]).cuda()
    res = ex.dispatch_local(x, r.router_mask.cuda(), 0)
    want = O.dispatch_local(c["hidden"].reshape(-1, c["H"]), r.router_mask, c["experts"], O.MIXTRAL_MOE_DENSE_ACT_DENSE)
    assert [t[2] for t in res] == [t[2] for t in want]
    for (o, l, e, hit), (wo, _, we, _) in zip(res, want):
        assert o.shape == wo.shape and o.dtype == x.dtype and o.device == x.device and l == 0
        hidden_close(o, wo, None, c["dtype"], f"expert {e}")
    assert [t[3] for t in res] == [0] * len(res)          # first touch: all misses
    res2 = ex.dispatch_local(x, r.router_mask.cuda(), 0)
    assert [t[3] for t in res2] == [1] * len(res2)        # now cached
    assert h.get_node_device(ids[res[0][2]]) == 0

    # whole block through the compat path (no moe_forward on the executor)
    class OnlyDispatch:
        def __init__(self, inner):
            self.dispatch_local = inner.dispatch_local
    cfg = types.SimpleNamespace(hidden_size=c["H"], intermediate_size=c["I"], num_local_experts=E, num_experts_per_tok=c["k"])
    blk = SyncMixtralSparseMoeBlock(cfg).to(torch.bfloat16).cuda()
    with torch.no_grad():
        blk.gate.weight.copy_(c["gate"])
    blk.expert_executor, blk.layer_id = OnlyDispatch(ex), 0
    with torch.no_grad():
        out, logits = blk(c["hidden"].cuda())
    same = (logits.cpu() == fx["router_logits"]).all(dim=-1) & ~fx["tied"]
    T = c["B"] * c["S"]
    hidden_close(out.reshape(T, -1)[same.cuda()], fx["out"].reshape(T, -1)[same], None, c["dtype"], "compat block")
    d.clear_expert_cache_counts()


def test_deepseek_block_golden(lib_built):
    from moe_infinity_b200.blocks import DeepseekMoEBlock
    from moe_infinity_b200.compat import DistributedExpertExecutor
    c, fx = load_case("deepseek_mini_bf16")
    eng = make_engine(c, "deepseek")
    cfg = types.SimpleNamespace(num_experts_per_tok=c["k"], n_routed_experts=c["E"], hidden_size=c["H"])
    blk = DeepseekMoEBlock(cfg).cuda()
    with torch.no_grad():
        blk.gate_weight.copy_(c["gate"].float())
    ex = DistributedExpertExecutor()
    ex.set_expert_dispatcher(eng)
    blk.expert_executor, blk.layer_id = ex, 0
    with torch.no_grad():
        out = blk(c["hidden"].cuda())
    T = c["B"] * c["S"]
    idx = eng.ws("topk_idx", T).cpu().long().sort(-1).values
    same = (idx == fx["topk_idx"].sort(-1).values).all(-1) & ~fx["tied"]
    assert same.float().mean() > 0.8
    o, r = out.reshape(T, -1)[same.cuda()].float().cpu(), fx["out"].reshape(T, -1)[same].float()
    eps = torch.finfo(torch.bfloat16).eps
    # gate scores come from a GPU fp32 GEMM here (CPU in the fixture): weights differ in the last fp32 bits
    assert torch.all((o - r).abs() <= 2 * eps * r.abs() + 2 * eps * r.pow(2).mean().sqrt())


def test_switch_block_golden(lib_built):
    from moe_infinity_b200.blocks import SyncSwitchTransformersSparseMLP
    from moe_infinity_b200.compat import DistributedExpertExecutor
    c, fx = load_case("switch_mini_bf16")
    eng = make_engine(c, "switch")
    cfg = types.SimpleNamespace(num_experts=c["E"], d_model=c["H"])
    blk = SyncSwitchTransformersSparseMLP(cfg).cuda()
    with torch.no_grad():
        blk.classifier.weight.copy_(c["gate"])
    ex = DistributedExpertExecutor()
    ex.set_expert_dispatcher(eng)
    blk.expert_executor, blk.layer_id = ex, 0
    with torch.no_grad():
        out, (logits, expert_index) = blk(c["hidden"].cuda())
    assert expert_index.shape == (c["B"], c["S"]) and out.shape == c["hidden"].shape
    stable = (logits.cpu().reshape(-1, c["E"]) - fx["router_logits"].reshape(-1, c["E"])).abs().max(-1).values < 1e-4
    assert torch.equal(expert_index.cpu().flatten()[stable], fx["expert_index"].flatten()[stable])
    hidden_close(out, fx["out"], None, c["dtype"], "switch block")


def test_dense_parameter_begin_is_static_placement(lib_built):
    """prefetch_handle.begin/end (model_offload.py:904-991 hooks): the dense tensor is uploaded on its first begin and stays
    resident -- later begins are free (same device storage), end releases nothing."""
    from moe_infinity_b200.compat import prefetch_handle
    h = prefetch_handle("/tmp/b2m_dense_unused", 0.5)
    w = torch.randn(64, 32)
    h.offload(w, 7)
    placeholder = torch.zeros(1)
    h.register(placeholder, 7)
    h.begin(0, placeholder)
    assert placeholder.is_cuda and torch.equal(placeholder.cpu(), w)
    p0 = placeholder.data_ptr()
    h.end(0, placeholder)
    h.begin(1, placeholder)
    assert placeholder.data_ptr() == p0          # no second upload
    other = torch.zeros(1)
    h.register(other, 7)                          # the same tensor id reached through another placeholder: same storage
    h.begin(2, other)
    assert other.data_ptr() == p0


def test_reference_compat_objects_with_experts_on_disk(lib_built, tmp_path):
    """Same call sequence as the reference's start-up (offload every tensor, register the experts by id), with
    `experts_on_disk=True`: nothing stays in host DRAM, every miss streams from the offload directory.  The
    per-expert outputs must be bit-identical to the host-DRAM handle's, through evictions (4 slots for 8 experts)."""
    from moe_infinity_b200 import _lib as L
    from moe_infinity_b200.compat import DistributedExpertExecutor, expert_dispatcher, prefetch_handle
    c, fx = load_case("mixtral_mini_bf16")
    E = c["E"]

    def build(handle):
        tid, ids = 0, {}
        for e in range(E):
            ids[e] = []
            for w in c["experts"][e]:
                handle.offload(w, tid)
                ids[e].append(tid)
                tid += 1
        d = expert_dispatcher(E, 1, L.DTYPE_BF16, L.EXPERT_MIXTRAL, 8, handle=handle, top_k=c["k"], max_tokens=64, num_slots=4)
        for e in range(E):
            d.register_expert(0, e, ids[e])
        ex = DistributedExpertExecutor()
        ex.set_expert_dispatcher(d)
        return ex, d

    ex_host, _ = build(prefetch_handle(str(tmp_path / "unused"), 0.5))
    h_disk = prefetch_handle(str(tmp_path / "offload"), 0.5, experts_on_disk=True)
    ex_disk, d_disk = build(h_disk)
    assert len(dict.keys(h_disk._tensors)) == 0                     # no expert tensor is held in host memory any more
    assert h_disk.is_tensor_offloaded(0)                             # ... but it is still known to the store
    r = O.mixtral_route(fx["router_logits"], c["k"], c["dtype"])
    x = c["hidden"].reshape(-1, c["H"]).cuda()
    mask = r.router_mask.cuda()
    for rep in range(3):
        for lo in (0, 4):                                             # two halves of the experts alternate -> evictions
            m = torch.zeros_like(mask)
            m[:, lo:lo + 4] = mask[:, lo:lo + 4]
            a = ex_host.dispatch_local(x, m, 0)
            b = ex_disk.dispatch_local(x, m, 0)
            assert [t[2:] for t in a] == [t[2:] for t in b]
            for (oa, _, e, _), (ob, _, _, _) in zip(a, b):
                assert torch.equal(oa, ob), f"rep {rep} expert {e}"
    st = d_disk.engine.stats()
    assert st["evictions"] > 0
    assert h_disk._readers and h_disk._readers[-1].stats()["bytes_read"] == st["misses"] * 3 * c["H"] * c["I"] * 2

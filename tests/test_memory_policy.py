"""CPU: host policy code -- tracer/predictor/prefetcher mirrors against the LITERAL reference classes
(/root/reference/moe_infinity/memory/*.py, dev container only) and the cache-policy oracle's stated rules."""
import contextlib
import os
import types

import numpy as np
import pytest
import torch

from moe_infinity_b200 import memory as M
from oracle.policy_oracle import CacheOracle

HAVE_REF = os.path.isdir("/root/reference/moe_infinity/memory")


@contextlib.contextmanager
def _cpu_only_torch():
    """The literal ExpertTracer allocates on cuda:0 (expert_tracer.py:33-35,104); run it on CPU."""
    zeros, to = torch.zeros, torch.Tensor.to

    def zeros_cpu(*a, **k):
        k.pop("device", None)
        return zeros(*a, **k)

    def to_cpu(self, *a, **k):
        if any(isinstance(x, str) and x == "cpu" for x in a):
            return self.clone()      # cuda:0 -> cpu is a copy in the real run; keep that (the caller mutates it)
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
        if not a and not k:
            return self
        return to(self, *a, **k)

    torch.zeros, torch.Tensor.to = zeros_cpu, to_cpu
    try:
        yield
    finally:
        torch.zeros, torch.Tensor.to = zeros, to


def _cfg(L, E):
    return types.SimpleNamespace(architectures=["MixtralForCausalLM"], num_hidden_layers=L, num_local_experts=E)


def _library(rng, n, L, E):
    lib = rng.integers(0, 6, size=(n, L, E)).astype(np.float32)
    lib[:, :, 0] += 1.0   # no all-zero rows
    return lib


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_predictor_matches_literal_reference(seed):
    import ref_loader
    ns = ref_loader.load()
    L, E, cap = 6, 8, 12
    rng = np.random.default_rng(seed)
    lib = _library(rng, 9, L, E)
    with _cpu_only_torch():
        ns.expert_tracer.ExpertTracer._instance = None
        rt = ns.expert_tracer.ExpertTracer(cap, _cfg(L, E))
        rt.trace_collection[:9] = torch.from_numpy(lib)
        rp = ns.expert_predictor.ExpertPredictor(_cfg(L, E))
        rp.add_tracer(rt)
        ot = M.ExpertTracer(cap, L, E)
        ot.load_trace(lib)
        op = M.ExpertPredictor(L, E)
        op.add_tracer(ot)
        rs, os_ = rt.create_entry(), ot.create_entry()
        for step in range(3):
            for layer in range(L):
                experts = rng.integers(0, E, size=(4, 2))
                a = rp.predict(rs, torch.from_numpy(experts), layer)
                b = op.predict(os_, experts, layer)
                np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9)
        np.testing.assert_array_equal(rt.get_entry(rs).matrix, ot.get_entry(os_).matrix)
        np.testing.assert_array_equal(rt.collection_access, ot.collection_access)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_prefetch_request_order_matches_literal_reference():
    import ref_loader
    ns = ref_loader.load()
    L, E = 5, 4
    rng = np.random.default_rng(3)
    matrix = rng.random((L, E)) * (rng.random((L, E)) > 0.3)

    class Rec:
        def __init__(self):
            self.cands, self.enq = None, []

        def replace_cache_candidates(self, ids):
            self.cands = list(ids)

        def get_node_default_device(self, ids):
            return 0

        def enqueue_prefetch(self, tid, gpu):
            self.enq.append(tid)

    tmap = {(l, e): 100 + l * E + e for l in range(L) for e in range(E)}
    import io, contextlib as cl
    with cl.redirect_stdout(io.StringIO()):
        rp = ns.expert_prefetcher.ExpertPrefetcher(_cfg(L, E))
    rp.expert_tensor_map = tmap
    r1 = Rec()
    rp.set_archer_engine(r1)
    rp.prefetch_experts(2, matrix)
    op = M.ExpertPrefetcher(L, E)
    op.expert_tensor_map = tmap
    r2 = Rec()
    op.set_archer_engine(r2)
    op.prefetch_experts(2, matrix)
    assert r1.cands == r2.cands and r1.enq == r2.enq
    assert [tmap[p] for p, _ in op.ordered_requests(2, matrix)] == r1.enq


def test_degenerate_empty_library_prefetches_everything_nearest_first():
    """with no trace library the predictor degenerates to 'all experts of later layers, nearest first'."""
    L, E = 4, 3
    tr = M.ExpertTracer(5, L, E)
    pr = M.ExpertPredictor(L, E)
    pr.add_tracer(tr)
    sid = tr.create_entry()
    m = pr.predict(sid, np.array([[0, 1]]), 1)
    assert np.all(m[0] == 0) and np.all(m[1:] > 0)
    pf = M.ExpertPrefetcher(L, E)
    reqs = pf.ordered_requests(1, m)
    assert [p[0] for p, _ in reqs] == [1] * E + [2] * E + [3] * E


def test_tracer_finish_entry_and_counts():
    tr = M.ExpertTracer(2, 3, 4)
    sid = tr.create_entry()
    tr.update_entry(sid, np.array([[1, 1], [2, 3]]), 0)
    assert tr.get_entry(sid).matrix[0].tolist() == [0, 2, 1, 1]
    tr.update_entry(sid, np.array([[0, 1]]), 2)
    assert tr.get_entry(sid).num_new_tokens == 1
    tr.finish_entry(sid)
    assert tr.trace_collection[0].sum() == 6 and tr.collection_access[0] == 1


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("current_layer", [0, 2, 3, 5])
def test_priority_score_matches_literal_reference(current_layer):
    import importlib
    import ref_loader
    ref_loader.load()
    ps = importlib.import_module("moe_infinity.memory.expert_priority_score")
    ent = importlib.import_module("moe_infinity.memory.expert_entry")
    L, E = 6, 4
    rng = np.random.default_rng(current_layer)
    dec = rng.integers(0, 4, size=(L, E)).astype(np.float64)
    dec[1] = 0                                           # an all-zero layer row
    freq = {(int(e), int(l)): float(rng.integers(0, 5)) for l in range(L) for e in range(E) if rng.random() < 0.6}
    entry = ent.ExpertTraceEntry("s", dec.copy(), 0, 0)
    ref_list = ps.priority_score(freq, set(), set(), entry, current_layer, L)
    ref_m = np.zeros((L, E))
    for ce in ref_list:
        ref_m[ce.layer_idx, ce.expert_idx] = ce.r
    ours = M.priority_score_matrix(freq, dec, current_layer, L)
    np.testing.assert_allclose(ours, ref_m, rtol=1e-12, atol=0)


# ---------------------------------------------------------------- cache policy oracle: the stated rules
def test_cache_oracle_lfu_eviction_and_tie_order():
    c = CacheOracle(num_layers=2, num_experts=4, num_slots=3, policy="slots")
    assert c.dispatch(0, [0, 1]) == [(0, False), (1, False)]
    assert c.dispatch(0, [0]) == [(0, True)]                 # visits: (0,0)=2, (0,1)=1
    assert c.dispatch(1, [2]) == [(2, False)]                # third slot
    assert c.dispatch(1, [3]) == [(3, False)]                # evicts min visits, ties -> expert-major scan
    assert c.evicted_log == [(0, 1)]                         # (0,1) and (1,2) both have 1 visit; expert 1 scanned first
    assert c.stats["misses"] == 4 and c.stats["hits"] == 1
    # the reference's byte budget is charged for hits too (expert_dispatcher.cpp:266): after 2 misses + 1 hit the budget
    # of 3 experts is used up, so the third miss already evicts although a slot is physically free
    # (behaviour confirmed on the reference's real engine: tests/golden/policy_ref_trace.json never holds 9 of its 9 experts)
    r = CacheOracle(num_layers=2, num_experts=4, num_slots=3, policy="reference")
    r.dispatch(0, [0, 1]); r.dispatch(0, [0]); r.dispatch(1, [2]); r.dispatch(1, [3])
    assert r.evicted_log == [(0, 1), (1, 2)] and sum(r.resident) == 2


def test_cache_oracle_waves_when_active_set_exceeds_slots():
    c = CacheOracle(1, 4, 2)
    c.dispatch(0, [0, 1])
    # 3 active experts, 2 slots: wave 1 = {1 (hit)} + nothing else fits without evicting a still-to-run expert...
    res = c.dispatch(0, [1, 2, 3])
    assert res == [(1, True), (2, False), (3, False)]
    assert c.stats["evictions"] == 2 and sum(c.resident) == 2
    c0 = CacheOracle(1, 4, 0)
    with pytest.raises(RuntimeError):
        c0.dispatch(0, [0])


def test_cache_oracle_prefetch_respects_protection():
    c = CacheOracle(2, 4, 2)
    c.dispatch(0, [0, 1])
    c.prefetch_hint([(1, 0), (1, 1), (1, 2)], [0.9, 0.5, 0.7])
    # last dispatch is in use -> nothing evictable -> no prefetch
    assert c.stats["prefetch_issued"] == 0
    c.dispatch(1, [3])                  # evicts (0,0) [visits equal, expert-major]
    c.prefetch_hint([(0, 0)], [1.0])    # (0,1) evictable (not protected, not in use)
    assert c.stats["prefetch_issued"] == 1 and c.evicted_log[-1] == (0, 1)
    c.prefetch_hint([(0, 0), (0, 2)], [1.0, 0.5])   # (0,0) resident+protected; (1,3) in use -> (0,2) dropped
    assert c.stats["prefetch_issued"] == 1
    assert c.dispatch(0, [0]) == [(0, True)] and c.stats["prefetch_useful"] == 1

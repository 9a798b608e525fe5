"""GPU: the device-side activation tracer / predictor (csrc/tracer.cu) against the host classes of
moe_infinity_b200/memory.py -- themselves held to the literal reference classes by tests/test_memory_policy.py.
Per layer call ONE kernel updates every sequence's trace matrix (expert_tracer.py:78-84), finds its nearest library trace
(:94-125) and writes the decayed prediction (expert_predictor.py:17-35); no host synchronisation happens in the layer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from moe_infinity_b200 import memory as M  # noqa: E402


def _engine(L, E, k, T):
    from moe_infinity_b200 import MoEEngine
    eng = MoEEngine(num_layers=L, num_experts=E, hidden=128, inter=128, top_k=k, dtype=torch.bfloat16, max_tokens=max(T, 16),
                    num_slots=L * E)
    g = torch.Generator().manual_seed(1)
    for l in range(L):
        for e in range(E):
            eng.load_expert(l, e, [(torch.randn(128, 128, generator=g) * 0.05).to(torch.bfloat16) for _ in range(3)])
    return eng


@pytest.mark.parametrize("L,E,k,B,S", [(6, 8, 2, 4, 1), (5, 16, 4, 3, 7)])
def test_device_tracer_matches_host_classes(lib_built, L, E, k, B, S):
    from moe_infinity_b200.engine import DeviceExpertTracer
    T = B * S
    eng = _engine(L, E, k, T)
    cap = 8          # capacity == loaded entries: an EMPTY library row has NaN distance and wins the argmin (checked below)
    dev = DeviceExpertTracer(eng, capacity=cap, max_seqs=B)
    rng = np.random.default_rng(L * 100 + E)
    # library: each entry prefers a different subset of experts per layer (clear nearest neighbours)
    lib = np.zeros((8, L, E), dtype=np.float32)
    for c in range(8):
        for l in range(L):
            hot = rng.choice(E, size=max(2, E // 4), replace=False)
            lib[c, l, hot] = rng.integers(5, 40, size=len(hot))
            lib[c, l] += rng.integers(0, 3, size=E)
    dev.load_trace(lib)
    host = M.ExpertTracer(cap, L, E)
    host.load_trace(lib)
    pred = M.ExpertPredictor(L, E)
    pred.add_tracer(host)
    seqs = [host.create_entry() for _ in range(B)]
    for b in range(B):
        dev.create_entry(b)
    syncs0 = eng.stats()["host_syncs"]
    x = torch.randn(T, 128, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()
    checked = 0
    for step in range(3):
        for l in range(L):
            # routing chosen by us (router logits in): sequence b follows library entry b with noise
            logits = torch.full((T, E), -4.0)
            for b in range(B):
                p = lib[b % 8, l] + 0.5
                for s_ in range(S):
                    logits[b * S + s_] = torch.log(torch.tensor(p / p.sum())) + 0.3 * torch.randn(E, generator=torch.Generator().manual_seed(100 * step + 10 * l + b * S + s_))
            eng.route(l, x, router_logits=logits.to(torch.bfloat16).cuda())
            dev.update_predict(l, num_seqs=B, seq_len=S)
            idx = eng.ws("topk_idx", T).cpu().numpy().reshape(B, S, k)
            want_hint = np.zeros((L, E), dtype=np.float64)
            hint_ok = True
            for b in range(B):
                m = pred.predict(seqs[b], idx[b], l)                      # host: update_entry + find_most_similar + decay
                np.testing.assert_array_equal(dev.entry(b), host.get_entry(seqs[b]).matrix.astype(np.float32))
                # nearest neighbour: identical whenever the host's margin between best and runner-up is clear
                cur = host.get_entry(seqs[b]).matrix
                d = _distances(host, cur, l)
                order = np.argsort(d)
                if d[order[1]] - d[order[0]] > 1e-4:
                    assert dev.winner(b) == int(order[0]), (step, l, b)
                    np.testing.assert_allclose(dev.prediction(b), m, rtol=1e-6, atol=1e-12)
                    checked += 1
                    want_hint += m
                else:
                    hint_ok = False          # near-tie between two library entries: the device may legitimately pick the other
            if hint_ok:
                np.testing.assert_allclose(dev.hint(), want_hint, rtol=1e-5, atol=1e-9)
    assert checked >= B * L                       # the comparison was not vacuous
    assert eng.stats()["host_syncs"] == syncs0    # nothing in the layer path synchronised with the host
    # finish_entry with a full library of persistent entries: nothing is replaced on either side
    before = dev.library(7).copy()
    dev.finish_entry(0)
    np.testing.assert_array_equal(dev.library(7), before)
    eng.close()
    # a library with empty rows: the reference's argmin returns the first NaN distance = the first empty row (Q8), on the
    # host class and on the device alike; finish_entry then fills exactly that row
    eng2 = _engine(L, E, k, T)
    dev2 = DeviceExpertTracer(eng2, capacity=cap + 3, max_seqs=B)
    dev2.load_trace(lib)
    host2 = M.ExpertTracer(cap + 3, L, E)
    host2.load_trace(lib)
    pred2 = M.ExpertPredictor(L, E)
    pred2.add_tracer(host2)
    sid = host2.create_entry()
    logits = torch.randn(T, E, generator=torch.Generator().manual_seed(5))
    eng2.route(1, x, router_logits=logits.to(torch.bfloat16).cuda())
    dev2.update_predict(1, num_seqs=B, seq_len=S)
    idx = eng2.ws("topk_idx", T).cpu().numpy().reshape(B, S, k)
    m = pred2.predict(sid, idx[0], 1)
    assert dev2.winner(0) == cap
    np.testing.assert_allclose(dev2.prediction(0), m, rtol=1e-6, atol=1e-12)
    dev2.finish_entry(0)
    host2.finish_entry(sid)
    np.testing.assert_array_equal(dev2.library(cap), host2.trace_collection[cap])
    assert dev2.access_counts()[cap] == 1
    eng2.close()


def _distances(tracer, matrix, layer_idx):
    """cos distance of find_most_similar (memory.py / expert_tracer.py:94-125) for every library entry."""
    lib = tracer.trace_collection.copy()
    lib[:, : (layer_idx + 1), :] = 1e-9
    with np.errstate(divide="ignore", invalid="ignore"):
        lib = lib / lib.sum(axis=2, keepdims=True)
        m = matrix.astype(np.float64).copy()
        m = (m / m.sum(axis=1, keepdims=True)).astype(np.float32)
    m = np.nan_to_num(m)
    num = (m[None] * lib).sum(axis=2)
    den = np.maximum(np.linalg.norm(m, axis=1), 1e-6)[None] * np.maximum(np.linalg.norm(lib, axis=2), 1e-6)
    with np.errstate(invalid="ignore"):
        d = 1 - (num / den).mean(axis=1)
    return np.where(np.isnan(d), -np.inf, d)

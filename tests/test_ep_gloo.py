"""CPU, world_size=2, gloo: the expert-parallel exchange schedule of moe_infinity_b200.ep.EPMoE
(rank->expert ownership, fixed-capacity buffers, counts all-gather, two all-to-alls) with a CPU stand-in for the
device steps.  The stand-in implements pack/regroup/ungroup/unpack exactly as csrc/ep.cu documents them, using
the oracle for routing and expert math, so the end-to-end result must equal the single-process oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import moe_oracle as O

H, I, E, K, T = 64, 128, 8, 2, 6
DT = torch.bfloat16


class CpuOps:
    """Test-only device backend (mirrors the C-ABI call sequence b2m_route/ep_pack/ep_regroup/...)."""

    def __init__(self, experts, gate, rank, world):
        self.experts, self.gate, self.rank, self.world = experts, gate, rank, world
        self.el = E // world

    def route(self, layer, x, router_logits=None):
        self.x = x
        logits = torch.nn.functional.linear(x, self.gate)
        self.r = O.mixtral_route(logits, K, x.dtype)
        rows, toks = [], []
        self.counts = torch.zeros(E, dtype=torch.int32)
        for e in range(E):
            idx = self.r.router_mask[:, e].nonzero().flatten()
            self.counts[e] = len(idx)
            toks += idx.tolist()
            rows.append(x[idx])
        self.xp = torch.cat(rows)
        self.perm_token = toks
        self.offsets = torch.cat([torch.zeros(1, dtype=torch.long), self.counts.long().cumsum(0)])

    def pack(self, world, rank, cap, T_, send_rows, send_counts=None):
        for r in range(world):
            send_rows[r, cap].view(torch.int32)[:E] = self.counts      # inline counts row (csrc/ep.cu)
            a, b = int(self.offsets[r * self.el]), int(self.offsets[(r + 1) * self.el])
            send_rows[r, : b - a] = self.xp[a:b]

    def regroup(self, world, rank, cap, T_total, recv_rows, recv_counts=None):
        recv_counts = torch.stack([recv_rows[s, cap].view(torch.int32)[:E] for s in range(world)])
        self.groups = []          # per local expert: list of (source, slot) in source-major order
        for le in range(self.el):
            e = rank * self.el + le
            g = []
            for s in range(world):
                pre = int(recv_counts[s, rank * self.el: e].sum())
                g += [(s, pre + j) for j in range(int(recv_counts[s, e]))]
            self.groups.append(g)
        self.recv_rows = recv_rows

    def run_experts(self, layer, T_total):
        self.outs = {}
        for le, g in enumerate(self.groups):
            if not g:
                continue
            xin = torch.stack([self.recv_rows[s, c] for s, c in g])
            y = O.expert_ffn(xin, self.experts[self.rank * self.el + le], O.MIXTRAL_MOE_DENSE_ACT_DENSE)
            for (s, c), row in zip(g, y):
                self.outs[(s, c)] = row

    def ungroup(self, world, rank, cap, ret_rows):
        for (s, c), row in self.outs.items():
            ret_rows[s, c] = row

    def unpack(self, world, rank, cap, T_, back_rows):
        ys = []
        for r in range(world):
            a, b = int(self.offsets[r * self.el]), int(self.offsets[(r + 1) * self.el])
            ys.append(back_rows[r, : b - a])
        self.y = torch.cat(ys)

    def combine(self, layer, x, out):
        final = torch.zeros_like(x)
        for e in range(E):
            a, b = int(self.offsets[e]), int(self.offsets[e + 1])
            if b > a:
                idx = self.r.router_mask[:, e]
                final[idx] += self.y[a:b] * self.r.routing_weights_mask[idx, e][:, None]
        if out is not None:
            out.copy_(final)
            return out
        return final


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moe_infinity_b200.ep import EPMoE, local_experts, owner_rank
    experts = O.make_experts(E, H, I, DT, seed=1, std=0.05)
    g = torch.Generator().manual_seed(2)
    gate = (torch.randn(E, H, generator=g) * 0.3).to(DT)
    gx = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(T, H, generator=gx).to(DT)
    assert local_experts(rank, E, world) == list(range(rank * E // world, (rank + 1) * E // world))
    assert all(owner_rank(e, E, world) == e // (E // world) for e in range(E))
    ep = EPMoE(CpuOps(experts, gate, rank, world), num_experts=E, hidden=H, top_k=K, T_local=T, dtype=DT,
               device="cpu")
    out = ep.forward(0, x)
    ref, _, _ = O.mixtral_block(x[None], gate, experts, K)
    ok = torch.equal(out, ref[0])
    # a second call with other tokens reuses the buffers (stale rows beyond the counts must not leak)
    x2 = torch.randn(T, H, generator=gx).to(DT)
    out2 = ep.forward(0, x2)
    ref2, _, _ = O.mixtral_block(x2[None], gate, experts, K)
    ok2 = torch.equal(out2, ref2[0])
    q.put((rank, bool(ok), bool(ok2)))
    dist.barrier()
    dist.destroy_process_group()


def test_ep_exchange_schedule_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True), (1, True, True)]

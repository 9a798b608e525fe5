"""Expert-parallel MoE layer over the GPUs of one box (BASELINE config 5): one process per GPU,
torch.distributed (NCCL over NVLink/NVSwitch) for the exchange, libb2m.so kernels for everything else.

Sharding: rank r owns experts [r*E/N, (r+1)*E/N) of every layer (contiguous blocks, so the expert-sorted
gathered rows of a rank are already contiguous per destination).  Per layer:
    route (local tokens) -> ep_pack -> all_to_all(rows + one counts row per peer, fixed capacity cap = T_local*k)
    -> ep_regroup -> grouped GEMMs on the received rows -> ep_ungroup -> all_to_all back -> ep_unpack -> combine.
Nothing in the layer synchronises with the host; capacity is static so no rank ever needs another rank's counts
on the CPU.  The reference has no live collective (README.md:18) -- it moves rows with `.to(device)`
inside one process (core/parallel/expert_dispatcher.cpp:283-285,403-405).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import Optional

import torch
import torch.distributed as dist


def owner_rank(expert: int, num_experts: int, world: int) -> int:
    return expert // (num_experts // world)


def local_experts(rank: int, num_experts: int, world: int):
    el = num_experts // world
    return list(range(rank * el, (rank + 1) * el))


class _EngineOps:
    """Device steps of one EP layer, implemented by the CUDA engine."""

    def __init__(self, eng):
        self.eng = eng

    def _s(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def route(self, layer, x, router_logits=None):
        return self.eng.route(layer, x, router_logits=router_logits)

    def pack(self, world, rank, cap, T, send_rows, send_counts=None):
        e = self.eng
        e._ck(e.lib.b2m_ep_pack(e._h, world, rank, cap, T, C.c_void_p(send_rows.data_ptr()),
                                C.c_void_p(send_counts.data_ptr() if send_counts is not None else 0), self._s()))

    def regroup(self, world, rank, cap, T_total, recv_rows, recv_counts=None):
        e = self.eng
        e._ck(e.lib.b2m_ep_regroup(e._h, world, rank, cap, T_total, C.c_void_p(recv_rows.data_ptr()),
                                   C.c_void_p(recv_counts.data_ptr() if recv_counts is not None else 0), self._s()))

    def run_experts(self, layer, T_total):
        self.eng.run_experts(layer, T_total)

    def ungroup(self, world, rank, cap, ret_rows):
        e = self.eng
        e._ck(e.lib.b2m_ep_ungroup(e._h, world, rank, cap, C.c_void_p(ret_rows.data_ptr()), self._s()))

    def unpack(self, world, rank, cap, T, back_rows):
        e = self.eng
        e._ck(e.lib.b2m_ep_unpack(e._h, world, rank, cap, T, C.c_void_p(back_rows.data_ptr()), self._s()))

    def combine(self, layer, x, out):
        return self.eng.combine(layer, x, out=out)

    # ---- peer-to-peer exchange (no collective call inside the layer)
    def p2p_setup(self, world, rank, cap, group):
        """Allocate this rank's exchange areas, swap CUDA IPC handles with the peers (once), map them."""
        e = self.eng
        h = (C.c_ubyte * 64)()
        e._ck(e.lib.b2m_ep_p2p_init(e._h, world, rank, cap, h))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h), group=group)
        for r, hb in enumerate(handles):
            buf = (C.c_ubyte * 64).from_buffer_copy(hb)
            e._ck(e.lib.b2m_ep_p2p_open(e._h, r, buf))
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def p2p_route(self, layer, x, router_logits=None):
        """route + dispatch fused (<= 256 tokens): gathered rows are stored straight into the owners' buffers."""
        e = self.eng
        x2 = e._check_x(x)
        rin, kind, rdt = e._router_args(router_logits, None)
        e._ck(e.lib.b2m_ep_p2p_route(e._h, layer, C.c_void_p(x2.data_ptr()),
                                     C.c_void_p(rin.data_ptr() if rin is not None else 0), kind, rdt, x2.shape[0],
                                     self._s()))

    def p2p_combine(self, layer, x, out):
        """collect + combine fused: the combine kernel reads the returned rows in place."""
        e = self.eng
        x2 = e._check_x(x)
        if out is None:
            out = torch.empty_like(x2)
        e._ck(e.lib.b2m_ep_p2p_combine(e._h, layer, C.c_void_p(x2.data_ptr()), x2.shape[0],
                                       C.c_void_p(out.data_ptr()), self._s()))
        return out

    def p2p_layer(self, layer, x, out, router_logits=None):
        """The whole expert-parallel layer in one C call (b2m_ep_p2p_layer): five kernels when a rank owns <= 8 experts."""
        e = self.eng
        x2 = e._check_x(x)
        if out is None:
            out = torch.empty_like(x2)
        rin, kind, rdt = e._router_args(router_logits, None)
        e._ck(e.lib.b2m_ep_p2p_layer(e._h, layer, C.c_void_p(x2.data_ptr()),
                                     C.c_void_p(rin.data_ptr() if rin is not None else 0), kind, rdt, x2.shape[0],
                                     C.c_void_p(out.data_ptr()), self._s()))
        return out

    def p2p_dispatch(self, T):
        e = self.eng
        e._ck(e.lib.b2m_ep_p2p_dispatch(e._h, T, self._s()))

    def p2p_regroup(self, T_total):
        e = self.eng
        e._ck(e.lib.b2m_ep_p2p_regroup(e._h, T_total, self._s()))

    def p2p_return(self):
        e = self.eng
        e._ck(e.lib.b2m_ep_p2p_return(e._h, self._s()))

    def p2p_collect(self, T):
        e = self.eng
        e._ck(e.lib.b2m_ep_p2p_collect(e._h, T, self._s()))


class EPMoE:
    """One expert-parallel MoE layer stack.  `ops` is the device backend (the CUDA engine in production; tests inject
    a CPU stand-in to exercise the exchange schedule under gloo)."""

    def __init__(self, ops, *, num_experts: int, hidden: int, top_k: int, T_local: int, dtype, device,
                 group: Optional[dist.ProcessGroup] = None, p2p: bool = False, fused: bool = True):
        self.ops = ops
        self.p2p = p2p
        self.fused = fused
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if num_experts % self.world:
            raise ValueError("num_experts must be divisible by the world size")
        self.E, self.H, self.k, self.T = num_experts, hidden, top_k, T_local
        self.cap = T_local * top_k
        self.T_total = self.world * T_local
        # inline layout: one extra row per peer segment carries counts[E] (int32), so each direction is ONE collective
        if 4 * num_experts > 2 * hidden:
            raise ValueError("inline counts need 4*E <= 2*H")
        if p2p:
            # rows are stored directly into the owners' buffers by the dispatch kernel (NVLink stores + flags)
            self.ops.p2p_setup(self.world, self.rank, self.cap, group)
            return
        shape = (self.world, self.cap + 1, hidden)
        self.send_rows = torch.zeros(shape, dtype=dtype, device=device)
        self.recv_rows = torch.zeros(shape, dtype=dtype, device=device)
        self.ret_rows = torch.zeros(shape, dtype=dtype, device=device)
        self.back_rows = torch.zeros(shape, dtype=dtype, device=device)

    def forward(self, layer: int, x: torch.Tensor, out: Optional[torch.Tensor] = None, router_logits=None):
        T = x.shape[0]
        if T != self.T:
            raise ValueError(f"EPMoE was sized for T_local={self.T}, got {T}")
        w, r, cap = self.world, self.rank, self.cap
        if self.p2p and T <= 256 and self.fused:
            if hasattr(self.ops, "p2p_layer"):
                return self.ops.p2p_layer(layer, x, out, router_logits)   # one C call; 5 kernels (direct mode) or the 7 below
            self.ops.p2p_route(layer, x, router_logits)          # gate/top-k + permute-and-dispatch (2 kernels)
            self.ops.p2p_regroup(self.T_total)
            self.ops.run_experts(layer, self.T_total)
            self.ops.p2p_return()
            return self.ops.p2p_combine(layer, x, out)            # collect-and-combine (1 kernel)
        self.ops.route(layer, x, router_logits)
        if self.p2p:
            self.ops.p2p_dispatch(T)
            self.ops.p2p_regroup(self.T_total)
            self.ops.run_experts(layer, self.T_total)
            self.ops.p2p_return()
            self.ops.p2p_collect(T)
            return self.ops.combine(layer, x, out)
        self.ops.pack(w, r, cap, T, self.send_rows)
        dist.all_to_all_single(self.recv_rows, self.send_rows, group=self.group)
        self.ops.regroup(w, r, cap, self.T_total, self.recv_rows)
        self.ops.run_experts(layer, self.T_total)
        self.ops.ungroup(w, r, cap, self.ret_rows)
        dist.all_to_all_single(self.back_rows, self.ret_rows, group=self.group)
        self.ops.unpack(w, r, cap, T, self.back_rows)
        return self.ops.combine(layer, x, out)


# --------------------------------------------------------------------------------------------- bench (N > 1)
def bench_ep(args, cfg, batch, metric, load_peaks, ClockSampler):
    from .engine import MoEEngine
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L, E, H, I, k = (args.layers or cfg["L"]), cfg["E"], cfg["H"], cfg["I"], cfg["k"]
    if E % world:
        raise SystemExit(f"--gpus {world} does not divide the {E} experts")
    T = batch
    dtype = torch.bfloat16
    mine = local_experts(rank, E, world)
    eng = MoEEngine(num_layers=L, num_experts=E, hidden=H, inter=I, top_k=k, dtype=dtype,
                    max_tokens=max(world * T, 16), num_slots=L * len(mine), device=local)
    torch.manual_seed(1000 + rank)

    def init_expert(view, l, e):
        # layer 0 is reproducible on every rank (per-expert generator) so that each rank can rebuild the whole layer for
        # the bit-equality check against the single-GPU engine below; deeper layers only need realistic values
        if l == 0:
            view.normal_(0.0, 0.02, generator=torch.Generator(device=dev).manual_seed(7000 + e))
        else:
            view.normal_(0.0, 0.02)
    for l in range(L):
        for e in mine:
            init_expert(eng.load_expert(l, e), l, e)
    g = torch.Generator(device=dev).manual_seed(0)           # gates are replicated: same seed on every rank
    for l in range(L):
        eng.set_gate(l, torch.randn(E, H, device=dev, generator=g) * 0.02)
    x_dev = torch.randn(L, T, H, device=dev).to(dtype)
    out_dev = torch.empty_like(x_dev)
    use_p2p = os.environ.get("B2M_EP_EXCHANGE", "p2p") == "p2p"
    ep = None
    if use_p2p:
        # CUDA IPC mapping of the peers' buffers can be unavailable (container/driver policy): agree on a fallback
        ok = torch.ones(1, device=dev)
        try:
            ep = EPMoE(_EngineOps(eng), num_experts=E, hidden=H, top_k=k, T_local=T, dtype=dtype, device=dev, p2p=True)
        except Exception as ex:  # pragma: no cover
            print(f"[rank {rank}] peer-to-peer exchange unavailable ({type(ex).__name__}: {ex}); using NCCL", flush=True)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            use_p2p, ep = False, None
    if ep is None:
        ep = EPMoE(_EngineOps(eng), num_experts=E, hidden=H, top_k=k, T_local=T, dtype=dtype, device=dev, p2p=False)

    def step():
        for l in range(L):
            ep.forward(l, x_dev[l], out=out_dev[l])

    # ---- parity at full size: this rank's tokens through the expert-parallel layer 0 vs a single-GPU engine that holds all E
    # experts of layer 0.  Same kernels, same rounding chain, experts combined in expert order; the only freedom is the
    # order of the split-K fp32 reductions of the down projection (atomics: not even the single-GPU engine repeats itself
    # bit for bit there, tests/test_gpu_fullsize.py), so: the tests' hidden-state bound (2 ulp of the value + 2 ulp of the rms,
    # tests/test_gpu_parity.py:hidden_close -- a one-ulp flip of one of the two summed expert outputs can exceed one ulp of
    # their sum) everywhere and >= 97 % of the elements identical
    full0 = MoEEngine(num_layers=1, num_experts=E, hidden=H, inter=I, top_k=k, dtype=dtype, max_tokens=max(T, 16),
                      num_slots=E, device=local)
    for e in range(E):
        init_expert(full0.load_expert(0, e), 0, e)
    full0.set_gate(0, eng._gates[0])
    par = torch.ones(2, device=dev)
    eps = torch.finfo(dtype).eps
    for it in range(2):
        xq = torch.randn(T, H, device=dev, generator=torch.Generator(device=dev).manual_seed(90 + 10 * rank + it)).to(dtype)
        a = full0.forward(0, xq).float()
        b = ep.forward(0, xq).float()
        torch.cuda.synchronize()
        ok = bool(((a - b).abs() <= 2 * eps * a.abs() + 2 * eps * a.pow(2).mean().sqrt()).all())
        frac = float((a == b).float().mean())
        par[0] = min(float(par[0]), 1.0 if (ok and frac >= 0.97) else 0.0)
        par[1] = min(float(par[1]), frac)
    dist.all_reduce(par, op=dist.ReduceOp.MIN)
    ep_parity = bool(par[0].item() == 1)
    ep_parity_frac = float(par[1].item())
    full0.close()
    del full0
    torch.cuda.empty_cache()

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    launches0 = eng.stats()["kernel_launches"]
    step()
    launches_per_step = eng.stats()["kernel_launches"] - launches0
    # one CUDA graph per 32-layer step (the exchange -- peer-to-peer kernels or NCCL all-to-alls -- is captured too);
    # fall back to eager launches
    eager_step = step
    timed = "eager launches"
    if not os.environ.get("B2M_EP_NO_GRAPH"):
        try:
            graph = torch.cuda.CUDAGraph()
            s_cap = torch.cuda.Stream()
            s_cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_cap):
                eager_step()
            torch.cuda.current_stream().wait_stream(s_cap)
            torch.cuda.synchronize()
            dist.barrier()
            with torch.cuda.graph(graph):
                eager_step()
            torch.cuda.synchronize()
            step = graph.replay
            timed = ("CUDA graph replay of the step (kernels incl. the peer-to-peer exchange)" if use_p2p
                     else "CUDA graph replay of the step (kernels + NCCL all-to-all)")
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        except Exception as ex:  # pragma: no cover
            step = eager_step
            timed = f"eager launches (graph capture failed: {type(ex).__name__})"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ms = []
    e0.record()
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        step_ms.append((a, b))
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    total_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    p50 = sorted(a.elapsed_time(b) for a, b in step_ms)[len(step_ms) // 2]
    clocks = sampler.stop() if rank == 0 else None

    timeline = None
    if os.environ.get("B2M_TIMELINE") == "1" and use_p2p:
        # device timestamps of the last replayed step: per layer, microseconds relative to the layer's first kernel
        buf = (C.c_uint64 * (16 * L))()
        try:
            eng._ck(eng.lib.b2m_timeline_read(eng._h, buf, L))
            rows = []
            for l in range(L):
                v = [int(buf[16 * l + i]) for i in range(16)]
                t0 = v[0]
                names = {"gate_end": 1, "permute_start": 2, "permute_end": 3, "k3_start": 4, "k3_flags": 5, "k3_end": 6,
                         "k4_start": 8, "k4_end": 10, "combine_start": 12, "combine_flags": 13, "combine_end": 14}
                rows.append({n: round((v[i] - t0) / 1e3, 2) for n, i in names.items()})
                rows[-1]["t0_ns"] = t0
            timeline = rows
        except Exception as ex:  # pragma: no cover
            timeline = f"unavailable: {ex}"
    # e2e: host buffers, copies inside the timed region
    x_host = x_dev.cpu().pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x_in = torch.empty_like(x_dev)

    def step_e2e():
        x_in.copy_(x_host, non_blocking=True)
        for l in range(L):
            ep.forward(l, x_in[l], out=out_dev[l])
        out_host.copy_(out_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_e2e()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    wall = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    e2e_ms = float(wall.item()) * 1e3

    gathered = [None] * world
    dist.all_gather_object(gathered, timeline)
    if rank == 0:
        peak, peak_src = load_peaks()
        ms_per_step = total_ms / args.steps
        el = len(mine)
        bytes_rank = L * el * 3 * H * I * 2       # every local expert is hit w.h.p. at T_total*k >= 32 assignments
        ach = bytes_rank / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": metric, "value": world * batch * args.steps / (total_ms * 1e-3), "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "p50_token_latency_ms": p50, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"Mixtral-8x7B MoE dispatch path, expert-parallel over {world} GPUs "
                                   f"({el} experts/GPU/layer), decode batch {batch} per GPU (global {world*batch}), "
                                   f"{L} layers, bf16 random-init, all local experts HBM-resident",
                       "global_batch": world * batch, "layers": L, "parallelism": f"ep{world}",
                       "exchange": ("fused peer-to-peer, 5 kernels/layer: the permute kernel stores rows + slot tags into the owners' "
                                    "receive areas over NVLink, the owners' GEMMs read them in place, the sources' combine reads the "
                                    "owners' fp32 outputs in place; st.release.sys/ld.acquire.sys epoch flags, no collective call") if use_p2p else
                                   "one fixed-capacity all_to_all_single (NCCL) each way; counts ride in the row buffer",
                       "l2": "inputs larger than L2 (each rank streams %.1f GB of weights per step)" % (bytes_rank / 1e9),
                       "timed_region": timed + ", max over ranks"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "peak_source": peak_src, "scope": "whole step, per rank", "traffic": None,
                         "algorithmic_bytes_per_rank_step": bytes_rank},
            "cpu_baseline": None,
            "e2e": {"value": world * batch * args.steps / (e2e_ms * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": x_host.numel() * 2 * world, "d2h_bytes_per_step": out_host.numel() * 2 * world,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches_per_step * args.steps * world), "launches_per_step": int(launches_per_step),
            "clocks": clocks, "ep_parity": ep_parity, "ep_parity_frac_bit_identical": ep_parity_frac,
            "timeline_us": gathered if any(g is not None for g in gathered) else None,
        }
        print(json.dumps(line), flush=True)
    # teardown: a captured graph holds NCCL work; destroying the process group under it can hang, so drop the graph,
    # drain, and leave without running communicator destructors
    step = eager_step
    graph = None
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    import sys
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)

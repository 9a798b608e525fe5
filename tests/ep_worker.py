"""torchrun worker for tests/test_gpu_ep.py: EP over WORLD_SIZE GPUs vs the same layer on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "moe-infinity_b200")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

from moe_infinity_b200 import MoEEngine
from moe_infinity_b200.ep import EPMoE, _EngineOps, local_experts
from oracle import moe_oracle as O


def run_case(rank, world, local, dev, *, H, I, E, K, T, L, deepseek):
    from moe_infinity_b200 import _lib as LB
    dt = torch.bfloat16
    et = O.DEEPSEEK_MOE_DENSE_ACT_DENSE if deepseek else O.MIXTRAL_MOE_DENSE_ACT_DENSE
    experts = [O.make_experts(E, H, I, dt, seed=10 + l, expert_type=et, std=0.05) for l in range(L)]
    g = torch.Generator().manual_seed(3)
    gdt = torch.float32 if deepseek else dt
    gates = [(torch.randn(E, H, generator=g) * 0.3).to(gdt) for _ in range(L)]
    kw = dict(num_layers=L, num_experts=E, hidden=H, inter=I, top_k=K, dtype=dt, device=local)
    if deepseek:
        kw.update(expert_type=LB.EXPERT_DEEPSEEK, router=LB.ROUTER_DEEPSEEK_GREEDY, routed_scaling_factor=2.0)
    full = MoEEngine(max_tokens=64, **kw)
    parts = [MoEEngine(max_tokens=world * T, num_slots=L * E // world, **kw) for _ in range(2)]
    for l in range(L):
        for e in range(E):
            full.load_expert(l, e, experts[l][e])
        full.set_gate(l, gates[l])
        for p in parts:
            for e in local_experts(rank, E, world):
                p.load_expert(l, e, experts[l][e])
            p.set_gate(l, gates[l])
    ep = EPMoE(_EngineOps(parts[0]), num_experts=E, hidden=H, top_k=K, T_local=T, dtype=dt, device=dev)
    ep2 = EPMoE(_EngineOps(parts[1]), num_experts=E, hidden=H, top_k=K, T_local=T, dtype=dt, device=dev, p2p=True,
                fused=(os.environ.get("B2M_EP_FUSED", "1") == "1"))
    gx = torch.Generator().manual_seed(50 + rank)
    bad = 0
    for it in range(5):
        for l in range(L):
            x = torch.randn(T, H, generator=gx).to(dt).cuda()
            a = full.forward(l, x).clone()
            b = ep.forward(l, x).clone()          # NCCL all-to-all exchange
            c = ep2.forward(l, x)                 # fused peer-to-peer exchange
            torch.cuda.synchronize()
            if not torch.equal(a, b) or not torch.equal(a, c):
                bad += 1
                print(f"rank {rank} it {it} layer {l} deepseek={deepseek}: nccl diff {(a.float()-b.float()).abs().max().item()} "
                      f"p2p diff {(a.float()-c.float()).abs().max().item()}", flush=True)
    return bad


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    bad = run_case(rank, world, local, dev, H=256, I=512, E=8, K=2, T=12, L=2, deepseek=False)
    bad += run_case(rank, world, local, dev, H=256, I=128, E=64, K=6, T=9, L=1, deepseek=True)
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("EP_WORKER_RESULT", "OK" if int(t.item()) == 0 else f"MISMATCH {int(t.item())}", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()

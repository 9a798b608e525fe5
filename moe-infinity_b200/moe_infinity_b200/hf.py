"""HuggingFace plug-in helper: run a `transformers` Mixtral model with its MoE blocks executed by the CUDA engine.

The reference achieves this by monkey-patching HF classes at load time (moe_infinity/runtime/model_offload.py:285-321:
`MixtralSparseMoeBlock -> SyncMixtralSparseMoeBlock`).  The installed transformers (5.x) stores the experts of a layer as
fused 3-D parameters (`MixtralExperts.gate_up_proj [E, 2I, H]`, `down_proj [E, H, I]`) and its block returns the hidden
states only, so this helper swaps `layer.mlp` for an adapter with that signature; the adapter keeps the router GEMM in
PyTorch (same logits as HF) and hands everything after it to `MoEEngine.forward`.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import MoEEngine


class B200MixtralMoeAdapter(nn.Module):
    """Drop-in for transformers>=5 `MixtralSparseMoeBlock`: forward(hidden[B,S,H]) -> hidden[B,S,H]."""

    def __init__(self, engine: MoEEngine, layer_id: int, gate_weight: torch.Tensor):
        super().__init__()
        self.engine = engine
        self.layer_id = layer_id
        self.gate_weight = nn.Parameter(gate_weight.detach().clone(), requires_grad=False)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        shape = hidden_states.shape
        x = hidden_states.reshape(-1, shape[-1])
        router_logits = F.linear(x, self.gate_weight)      # same GEMM HF's MixtralTopKRouter runs
        out = self.engine.forward(self.layer_id, x, router_logits=router_logits)
        return out.reshape(shape)


def patch_mixtral(model, max_tokens: int = 4096, num_slots: int = 0, device_memory_ratio: float = 0.0,
                  resident: bool = True, engine: Optional[MoEEngine] = None) -> MoEEngine:
    """Move every expert of `model` (a transformers MixtralForCausalLM / MixtralModel on a CUDA device) into a
    MoEEngine and replace the MoE blocks.  `resident=False` registers experts as pinned host blobs (offload mode,
    budget from num_slots / device_memory_ratio) instead of loading all of them into HBM."""
    cfg = model.config
    layers = model.model.layers if hasattr(model, "model") else model.layers
    dev = next(model.parameters()).device
    dtype = next(model.parameters()).dtype
    L, E, H, I, k = len(layers), cfg.num_local_experts, cfg.hidden_size, cfg.intermediate_size, cfg.num_experts_per_tok
    if engine is None:
        engine = MoEEngine(num_layers=L, num_experts=E, hidden=H, inter=I, top_k=k, dtype=dtype, max_tokens=max_tokens,
                           num_slots=(L * E if resident and not num_slots else num_slots),
                           device_memory_ratio=device_memory_ratio, device=dev.index or 0)
    for l, layer in enumerate(layers):
        moe = layer.mlp
        gu, dn = moe.experts.gate_up_proj.data, moe.experts.down_proj.data     # [E,2I,H], [E,H,I]
        for e in range(E):
            w1, w3, w2 = gu[e, :I], gu[e, I:], dn[e]                            # reference order: w1 | w2 | w3
            if resident:
                engine.load_expert(l, e, [w1, w2, w3])
            else:
                engine.register_expert(l, e, [w1, w2, w3])
        layer.mlp = B200MixtralMoeAdapter(engine, l, moe.gate.weight.data)
    return engine

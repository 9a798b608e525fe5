"""Seeded expert-FFN cases shared by make_expert_golden.py (generator) and tests/test_oracle_expert_ref.py."""
from __future__ import annotations

import torch

# expert_type ints: core/parallel/expert_module.h:13-18; dtype ints :20-23
DT = {0: torch.bfloat16, 1: torch.float32, 2: torch.float16}
N_TENSORS = {0: 2, 1: 3, 2: 4, 3: 4, 4: 3, 5: 3}

# name: (expert_type, dtype_int, H, I, n_rows, seed)
CASES = {
    "switch_relu_f32": (0, 1, 64, 160, 7, 101),
    "switch_relu_bf16": (0, 0, 64, 160, 7, 102),
    "switch_gated_gelu_bf16": (1, 0, 96, 128, 5, 103),
    "switch_gated_gelu_f32": (1, 1, 96, 128, 5, 104),
    "nllb_bias_relu_bf16": (2, 0, 64, 192, 9, 105),
    "nllb_bias_relu_f16": (2, 2, 64, 192, 9, 106),
    "nllb_bias_relu_f32": (2, 1, 64, 192, 9, 107),
    "fsgpt_bias_relu_bf16": (3, 0, 128, 64, 3, 108),
    "fsgpt_bias_relu_f32": (3, 1, 128, 64, 3, 109),
    "mixtral_swiglu_bf16": (4, 0, 128, 256, 11, 110),
    "mixtral_swiglu_f16": (4, 2, 128, 256, 11, 111),
    "mixtral_swiglu_f32": (4, 1, 128, 256, 11, 112),
    "mixtral_one_row_bf16": (4, 0, 128, 256, 1, 113),
    "deepseek_swiglu_bf16": (5, 0, 128, 96, 6, 114),
    "deepseek_swiglu_f16": (5, 2, 128, 96, 6, 115),
}


def make_case(name):
    """-> (expert_type, dtype_int, [tensors in the reference's tensor-id order], x[n,H])"""
    et, di, H, I, n, seed = CASES[name]
    dt = DT[di]
    g = torch.Generator().manual_seed(seed)
    r = lambda *shape, std=0.08: (torch.randn(*shape, generator=g) * std).to(dt)  # noqa: E731
    if et == 0:
        ws = [r(I, H), r(H, I)]                       # wi, wo
    elif et in (1, 4, 5):
        # type 1: wi_0, wi_1, wo   type 4: w1, w2, w3   type 5: gate, up, down
        ws = [r(I, H), r(H, I), r(I, H)] if et == 4 else [r(I, H), r(I, H), r(H, I)]
    else:
        ws = [r(I, H), r(I, std=0.3), r(H, I), r(H, std=0.3)]   # fc1, fc1_bias, fc2, fc2_bias
    x = torch.randn(n, H, generator=g).to(dt)
    return et, di, ws, x

"""Generate tests/golden/expert_ffn_ref.pt from the reference's OWN expert modules (dev container only).

Run:  python tests/golden/make_expert_golden.py
Needs oracle/_ref/ref_expert_module.so (= /root/reference/core/parallel/expert_module.cpp compiled as-is, see
oracle/ref_build/Makefile).  Stores, per case, the input rows, a checksum of the seeded weights and the output of the
real `<Type>MoEDenseActDense::forward`; weights are regenerated from the seed (tests/golden/expert_cases.py).
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_module  # noqa: E402
import expert_cases as C  # noqa: E402


def main():
    torch.set_num_threads(1)      # single-thread ATen: the accumulation order is then fixed for these sizes
    ref_module.build(verbose=True)
    R = ref_module.load()
    assert R is not None, "oracle/_ref/ref_expert_module.so missing"
    out = {}
    for name in C.CASES:
        et, di, ws, x = C.make_case(name)
        y = R.expert_forward(et, di, ws, x)
        out[name] = dict(x=x, y=y, wsum=sum(float(w.double().abs().sum()) for w in ws))
        print(f"{name:28s} y {tuple(y.shape)} {y.dtype} |y|mean {y.float().abs().mean():.4f}")
    torch.save(out, os.path.join(HERE, "expert_ffn_ref.pt"))


if __name__ == "__main__":
    main()

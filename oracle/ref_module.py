"""Loader for oracle/_ref/ref_expert_module.so  --  TEST INFRASTRUCTURE ONLY.

The .so is the reference's own core/parallel/expert_module.cpp compiled where it lies (oracle/ref_build/Makefile)
plus a 60-line harness; it exposes `expert_forward(expert_type, dtype, tensors, x)` = the real
`<Type>MoEDenseActDense::forward` of the reference on CPU.  Used to pin oracle.moe_oracle.expert_ffn (D1-D3 of
SURVEY §8a) and to generate tests/golden/expert_ffn_ref.pt.  Never imported by the product.
"""
from __future__ import annotations

import importlib.util
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "ref_expert_module.so")


def build(reference_root: str = "/root/reference", verbose: bool = False) -> str | None:
    """Compile the reference TU + harness if the reference tree is present (dev container); returns the .so path."""
    src = os.path.join(reference_root, "core", "parallel", "expert_module.cpp")
    if not os.path.exists(src):
        return SO if os.path.exists(SO) else None
    deps = [src, os.path.join(reference_root, "core", "aio", "archer_tensor_index.cpp"),
            os.path.join(HERE, "ref_build", "ref_expert_harness.cpp"), os.path.join(HERE, "ref_build", "Makefile")]
    if os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(d) for d in deps):
        return SO
    r = subprocess.run(["make", "-C", os.path.join(HERE, "ref_build"), f"REF={reference_root}"],
                       capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building oracle/_ref failed:\n" + (r.stderr or ""))
    return SO


def load():
    """The compiled reference module, or None when it was never built (e.g. a checkout without /root/reference)."""
    if not os.path.exists(SO):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("ref_expert_module", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

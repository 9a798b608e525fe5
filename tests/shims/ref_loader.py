"""Load the *literal* reference hot-path files from /root/reference on CPU (dev container only).

The reference package cannot be imported as is (no `accelerate`, transformers 5.5 vs the
4.x names it uses, ExpertTracer allocating on cuda:0) -- SURVEY.md §8(c).  Four small
shims make the hot-path files load unmodified:
  1. stub packages `moe_infinity{,.models,.memory,.distributed,.utils}` whose __path__ points
     at the reference directories (bypasses the __init__ import chain);
  2. a fake `accelerate` exposing the three names the reference touches;
  3. a 4.x-style `MixtralBlockSparseTop2MLP` (w1,w2,w3 nn.Linear) injected into
     transformers.models.mixtral.modeling_mixtral;
  4. transformers.utils.import_utils.is_torch_fx_available = lambda: False.
Nothing here is copied from the reference; the files are executed from where they lie.
This module is test infrastructure and is never imported by the product package.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("B2M_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "moe_infinity", "models"))


_loaded = {}


def _stub_pkg(name: str, path: str):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the literal reference classes."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present")
    import torch
    import torch.nn as nn
    import transformers  # noqa: F401  (must be imported before the fake accelerate is installed)
    import transformers.models.mixtral.modeling_mixtral as mm
    import transformers.utils.import_utils as iu

    # shim 2: fake accelerate (only if the real one is missing)
    try:
        import accelerate  # noqa: F401
    except Exception:
        acc = types.ModuleType("accelerate")
        acc.__path__ = []
        acc_utils = types.ModuleType("accelerate.utils")
        acc_utils.__path__ = []
        acc_ver = types.ModuleType("accelerate.utils.versions")
        acc_ver.is_torch_version = lambda op, v: True
        acc_const = types.ModuleType("accelerate.utils.constants")
        acc_const.SAFE_WEIGHTS_NAME = "model.safetensors"
        acc_const.WEIGHTS_NAME = "pytorch_model.bin"
        sys.modules.update({"accelerate": acc, "accelerate.utils": acc_utils,
                            "accelerate.utils.versions": acc_ver,
                            "accelerate.utils.constants": acc_const})

    # shim 3: 4.x expert MLP (names and order w1, w2, w3 as in HF 4.x MixtralBlockSparseTop2MLP)
    if not hasattr(mm, "MixtralBlockSparseTop2MLP"):
        class MixtralBlockSparseTop2MLP(nn.Module):
            def __init__(self, config):
                super().__init__()
                self.ffn_dim = config.intermediate_size
                self.hidden_dim = config.hidden_size
                self.w1 = nn.Linear(self.hidden_dim, self.ffn_dim, bias=False)
                self.w2 = nn.Linear(self.ffn_dim, self.hidden_dim, bias=False)
                self.w3 = nn.Linear(self.hidden_dim, self.ffn_dim, bias=False)
                self.act_fn = nn.SiLU()

            def forward(self, hidden_states):
                return self.w2(self.act_fn(self.w1(hidden_states)) * self.w3(hidden_states))
        mm.MixtralBlockSparseTop2MLP = MixtralBlockSparseTop2MLP

    # shim 4
    if not hasattr(iu, "is_torch_fx_available"):
        iu.is_torch_fx_available = lambda: False

    # shim 1: stub packages
    base = os.path.join(REF_ROOT, "moe_infinity")
    _stub_pkg("moe_infinity", base)
    for sub in ("models", "memory", "distributed", "utils", "common"):
        _stub_pkg(f"moe_infinity.{sub}", os.path.join(base, sub))
    # moe_infinity.utils exports ArcherConfig + hf_config parsers via its __init__; load literally
    cfg = importlib.import_module("moe_infinity.utils.config")
    hf = importlib.import_module("moe_infinity.utils.hf_config")
    u = sys.modules["moe_infinity.utils"]
    u.ArcherConfig = cfg.ArcherConfig
    for n in ("parse_moe_param", "parse_expert_id", "parse_expert_dtype", "parse_expert_type"):
        if hasattr(hf, n):
            setattr(u, n, getattr(hf, n))

    ns = types.SimpleNamespace()
    ns.mixtral = importlib.import_module("moe_infinity.models.mixtral")
    ns.deepseek = importlib.import_module("moe_infinity.models.deepseek")
    ns.modeling_deepseek = importlib.import_module("moe_infinity.models.modeling_deepseek.modeling_deepseek")
    ns.expert_executor = importlib.import_module("moe_infinity.distributed.expert_executor")
    try:
        ns.expert_predictor = importlib.import_module("moe_infinity.memory.expert_predictor")
        ns.expert_prefetcher = importlib.import_module("moe_infinity.memory.expert_prefetcher")
        ns.expert_tracer = importlib.import_module("moe_infinity.memory.expert_tracer")
    except Exception as e:  # pragma: no cover - memory/* are optional for the block tests
        ns.memory_import_error = e
    _loaded["ns"] = ns
    return ns

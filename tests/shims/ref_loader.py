"""Load the *literal* reference hot-path files: from /root/reference (dev container) or, where that does not exist (the
GPU box), from their byte-compiled form under oracle/_ref/pyref/ (built by oracle/ref_build/stage_pyref.py).

The reference package cannot be imported as is (no `accelerate`, transformers 5.5 vs the
4.x names it uses, ExpertTracer allocating on cuda:0) -- SURVEY.md §8(c).  Four small
shims make the hot-path files load unmodified:
  1. stub packages `moe_infinity{,.models,.memory,.distributed,.utils}` whose __path__ points
     at the reference directories (bypasses the __init__ import chain);
  2. a fake `accelerate` exposing the three names the reference touches;
  3. a 4.x-style `MixtralBlockSparseTop2MLP` (w1,w2,w3 nn.Linear) injected into
     transformers.models.mixtral.modeling_mixtral;
  4. transformers.utils.import_utils.is_torch_fx_available = lambda: False;
  5. a 4.x-order `SwitchTransformersTop1Router` (returns `(expert_mask, router_probs, router_logits)`, the tuple
     moe_infinity/models/switch_transformers.py:76 unpacks) installed in HF's modeling_switch_transformers: HF 5.5's
     router returns `(probs, index, logits)` and computes the capacity cumsum over a size-1 axis.  The shim restates
     the forward of transformers 4.37-4.4x (the version range the reference pins, requirements.txt:19).
Nothing here is copied from the reference; the files are executed from where they lie.
This module is test infrastructure and is never imported by the product package.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_CANDIDATES = [os.environ.get("B2M_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "oracle", "_ref", "pyref")]
REF_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "moe_infinity", "models"))), "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "moe_infinity", "models"))


def is_source_tree() -> bool:
    """True when REF_ROOT holds the reference's .py sources (dev container), False for the byte-compiled staging."""
    return os.path.exists(os.path.join(REF_ROOT, "moe_infinity", "models", "mixtral.py"))


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

    # shim 5: HF 4.x router output order and capacity semantics for the literal Switch block
    import transformers.models.switch_transformers.modeling_switch_transformers as ms
    if not getattr(ms.SwitchTransformersTop1Router, "_b2m_4x_order", False):
        _Base = ms.SwitchTransformersTop1Router

        class SwitchTransformersTop1Router(_Base):   # same name, same constructor
            _b2m_4x_order = True

            def forward(self, hidden_states):
                self.input_dtype = hidden_states.dtype
                hidden_states = hidden_states.to(self.dtype)
                self.classifier = self.classifier.to(self.dtype)
                router_logits = self.classifier(hidden_states)
                router_probs = nn.functional.softmax(router_logits, dim=-1, dtype=self.dtype).to(self.input_dtype)
                expert_index = torch.argmax(router_probs, dim=-1)
                expert_index = torch.nn.functional.one_hot(expert_index, num_classes=self.num_experts)
                token_priority = torch.cumsum(expert_index, dim=-2)          # per batch row, over the sequence
                expert_capacity_mask = token_priority <= self.expert_capacity
                expert_index = expert_index * expert_capacity_mask
                router_probs = torch.max(router_probs, dim=-1).values.unsqueeze(-1)
                return expert_index, router_probs, router_logits
        ms.SwitchTransformersTop1Router = SwitchTransformersTop1Router

    # shim 6: HF 4.x contract of the NLLB router for the literal NLLB block (nllb_moe.py:53 unpacks two values; 4.x flattens
    # the tokens before the classifier and returns (top_1_mask, router_probs); HF 5.x adds the logits as a third value)
    try:
        import transformers.models.nllb_moe.modeling_nllb_moe as mn
        if not getattr(mn.NllbMoeTop2Router, "_b2m_4x_contract", False):
            _NBase = mn.NllbMoeTop2Router

            class NllbMoeTop2Router(_NBase):   # same name, same constructor
                _b2m_4x_contract = True

                def forward(self, hidden_states, padding_mask=None):
                    self.input_dtype = hidden_states.dtype
                    b, s_, h = hidden_states.shape
                    hidden_states = hidden_states.reshape(b * s_, h).to(self.dtype)
                    self._cast_classifier()
                    router_logits = self.classifier(hidden_states)
                    top_1_mask, router_probs = self.route_tokens(router_logits, self.input_dtype, padding_mask)
                    return top_1_mask, router_probs
            mn.NllbMoeTop2Router = NllbMoeTop2Router
    except Exception:  # pragma: no cover - the NLLB block is optional
        pass

    # byte-compiled staging only: HF's docstring decorators call inspect.getsource() on the decorated forward(); there is
    # no source text on the GPU box, so fall back to the default indentation (affects generated docstrings only)
    if not is_source_tree():
        import transformers.utils.doc as hfdoc
        _orig_indent = hfdoc.get_docstring_indentation_level

        def _indent_or_default(fn):
            try:
                return _orig_indent(fn)
            except OSError:
                return 4
        hfdoc.get_docstring_indentation_level = _indent_or_default

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
        m = sys.modules["moe_infinity.memory"]
        for n, mod in (("ExpertPredictor", ns.expert_predictor), ("ExpertPrefetcher", ns.expert_prefetcher),
                       ("ExpertTracer", ns.expert_tracer)):
            if not hasattr(m, n):
                setattr(m, n, getattr(mod, n))
    except Exception as e:  # pragma: no cover - memory/* are optional for the block tests
        ns.memory_import_error = e
    try:
        ns.switch = importlib.import_module("moe_infinity.models.switch_transformers")   # needs memory.ExpertPredictor
    except Exception as e:  # pragma: no cover
        ns.switch_import_error = e
    try:
        ns.nllb = importlib.import_module("moe_infinity.models.nllb_moe")
    except Exception as e:  # pragma: no cover
        ns.nllb_import_error = e
    _loaded["ns"] = ns
    return ns

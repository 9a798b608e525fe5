"""Reference-compatible host objects: `prefetch_handle`, `expert_dispatcher`, `DistributedExpertExecutor`.

Same class names, method names, argument meaning and result contracts as the reference's pybind module
`moe_infinity.ops.prefetch.prefetch_op` (core/python/py_archer_prefetch.cpp:10-92) and
moe_infinity/distributed/expert_executor.py:19-58, restricted to what the MoE-block hot path calls
(SURVEY.md §8 b.2).  Everything numerical is forwarded to libb2m.so through MoEEngine.

Differences, on purpose:
  * `offload` keeps tensors in host DRAM; with `persistent=True` it also maintains the reference's on-disk store
    (store.py reads and writes `archer_index` / `archer_param_0` in the reference's format) -- the asynchronous disk
    reader itself (core/aio) is out of scope;
  * errors raise RuntimeError/B2MError instead of aborting the process;
  * `wait_expert` returns experts in ascending expert id (the reference returns completion order, Q1).

Construction is signature-compatible with the reference: `prefetch_handle(prefix, ratio)` and
`expert_dispatcher(num_experts, num_layers, dtype, expert_type, num_threads)` (py_archer_prefetch.cpp:12,85), so
moe_infinity/runtime/model_offload.py:143-145,471-477 runs unchanged.  The dispatcher finds its handle the way the
reference does -- through process-wide state (the reference's kTopologyHandle / kArcherTensorHandle singletons,
archer_prefetch_handle.cpp:18-28): the most recently constructed prefetch_handle -- and learns top_k from the router
masks it is handed (the reference has no such notion: every set mask bit runs, expert_dispatcher.cpp:274-285).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .engine import MoEEngine
from .store import ArcherTensorStore

_INT2DTYPE = {L.DTYPE_BF16: torch.bfloat16, L.DTYPE_F32: torch.float32, L.DTYPE_F16: torch.float16}
_ROUTER_OF_TYPE = {L.EXPERT_MIXTRAL: L.ROUTER_MIXTRAL, L.EXPERT_DEEPSEEK: L.ROUTER_DEEPSEEK_GREEDY,
                   L.EXPERT_SWITCH: L.ROUTER_SWITCH_TOP1, L.EXPERT_SWITCH_GATED: L.ROUTER_SWITCH_TOP1,
                   # NLLB / FSGPT blocks route in Python (top-2 softmax) and reach the experts through dispatch_local
                   L.EXPERT_NLLB: L.ROUTER_MIXTRAL, L.EXPERT_FSGPT: L.ROUTER_MIXTRAL}


class _LazyTensors(dict):
    """id -> host tensor; ids that only exist in the on-disk store are read on first use."""

    def __init__(self, store: Optional[ArcherTensorStore]):
        super().__init__()
        self._store = store

    def __missing__(self, tensor_id):
        if self._store is None or tensor_id not in self._store:
            raise KeyError(tensor_id)
        t = self._store.read_tensor(tensor_id)
        self[tensor_id] = t
        return t

    def __contains__(self, tensor_id):
        return dict.__contains__(self, tensor_id) or (self._store is not None and tensor_id in self._store)


_CURRENT_HANDLE: Optional["prefetch_handle"] = None     # the reference keeps exactly one per process (global singletons)


class prefetch_handle:  # noqa: N801  (reference spelling)
    """py_archer_prefetch.cpp:11-80.  Host-DRAM tensor store + the residency/prefetch façade."""

    def __init__(self, prefix: str, device_memory_ratio: float, persistent: bool = False, experts_on_disk: bool = False):
        global _CURRENT_HANDLE
        _CURRENT_HANDLE = self
        self.prefix = prefix
        self.device_memory_ratio = float(device_memory_ratio)
        # experts_on_disk=True (implies persistent): expert tensors are NOT kept in host DRAM once registered; a cache miss
        # streams them from the offload directory through the native reader (disk -> pinned staging ring -> HBM slot).  For
        # models larger than host memory; the reference instead pins every expert in host DRAM at start-up
        # (model_topology.cpp:517-537).
        self.experts_on_disk = bool(experts_on_disk)
        persistent = persistent or self.experts_on_disk
        self._readers: List = []
        # persistent=True: `prefix` is an offload directory in the reference's on-disk format (store.py): `offload`
        # also writes there and a directory left by an earlier run (of this package or of the reference) is reused,
        # tensors being read back on first use.  Default: host DRAM only, nothing touches the disk.
        self._store = ArcherTensorStore(prefix) if persistent else None
        self._tensors: Dict[int, torch.Tensor] = _LazyTensors(self._store)     # id -> host tensor  (offload)
        self._params: Dict[int, torch.Tensor] = {}      # id -> live Parameter.data placeholder (register)
        self._ptr2id: Dict[int, int] = {}
        self._node_of: Dict[int, Tuple[int, int]] = {}  # tensor id -> (layer, expert)
        self._topology = None
        self._dispatcher: Optional["expert_dispatcher"] = None

    # ---- tensor store (prefetch_handle.offload / register / is_tensor_*)
    def offload(self, tensor: torch.Tensor, tensor_id: int):
        self._tensors[int(tensor_id)] = tensor.detach().to("cpu").contiguous()
        if self._store is not None:
            self._store.store_tensor(int(tensor_id), tensor, flush=False)

    def flush(self):
        """Write the on-disk index once (the reference rewrites it inside every `offload`, O(n^2) in the tensor count)."""
        if self._store is not None:
            self._store.flush()

    def _reader_for(self, tensor_ids: Sequence[int]):
        """The native store reader that knows `tensor_ids` (its index is read when it is opened: tensors offloaded later
        need a new one; earlier readers stay open because contexts hold on to them)."""
        from .store import NativeStoreReader
        if self._readers:
            try:
                self._readers[-1].blob_bytes(tensor_ids)
                return self._readers[-1]
            except ValueError:
                pass
        self.flush()
        self._readers.append(NativeStoreReader(self.prefix))
        return self._readers[-1]

    def _shape_of(self, tensor_id: int):
        if dict.__contains__(self._tensors, tensor_id):
            return tuple(self._tensors[tensor_id].shape)
        return tuple(self._store.index[int(tensor_id)].shape)

    def register(self, tensor: torch.Tensor, tensor_id: int):
        self._params[int(tensor_id)] = tensor
        self._ptr2id[tensor.data_ptr()] = int(tensor_id)

    def update_tensor_map(self, old_ptr: int, new_ptr: int):
        if old_ptr in self._ptr2id:
            self._ptr2id[new_ptr] = self._ptr2id.pop(old_ptr)

    def is_tensor_offloaded(self, tensor_id: int) -> bool:
        return int(tensor_id) in self._tensors

    def is_tensor_index_initialized(self) -> bool:
        # archer_tensor_handle.h:38: true iff an index file was found when the store was opened; without a persistent
        # store the caller re-offloads from the checkpoint (model_offload.py:351)
        return self._store is not None and self._store.is_initialized()

    def set_topology(self, topology):
        """[(name, [[ids...], ...]), ...]; a stage with >1 id-lists is an expert stage (model_topology.cpp:417-452)."""
        self._topology = topology

    def get_node_default_device(self, tensor_ids: Sequence[int]) -> int:
        return self._dispatcher.engine.device.index if self._dispatcher and self._dispatcher.engine else 0

    def get_node_device(self, tensor_ids: Sequence[int]) -> int:
        le = self._node_of.get(int(tensor_ids[0]))
        if le is None or self._dispatcher is None or self._dispatcher.engine is None:
            return -1
        return self._dispatcher.engine.device.index if self._dispatcher.engine.is_resident(*le) else -1

    # ---- dense-parameter path (begin/end hooks, model_offload.py:904-991 -> archer_prefetch_handle.cpp:83-180): static
    # placement -- a dense tensor is uploaded ONCE, on its first `begin`, and stays resident; later begins only find it already
    # in place (the reference re-points .data on every begin and puts a 1-element placeholder back on every end)
    def begin(self, request_id: int, tensor: torch.Tensor):
        ptr = tensor.data_ptr()
        tid = self._ptr2id.get(ptr)
        if tid is None or tid not in self._tensors:
            return
        dev_t = self._dense_dev.get(tid) if hasattr(self, "_dense_dev") else None
        if dev_t is None:
            if not hasattr(self, "_dense_dev"):
                self._dense_dev = {}
            dev = torch.device("cuda", self.get_node_default_device([tid]))
            dev_t = self._tensors[tid].to(dev)
            self._dense_dev[tid] = dev_t
        if ptr != dev_t.data_ptr():
            tensor.data = dev_t
            self._ptr2id[dev_t.data_ptr()] = tid

    def end(self, request_id: int, tensor: torch.Tensor):
        return None          # stays resident: nothing to release

    def fetch_tensors(self, request_id: int, tensor_ids: Sequence[int]):
        return None

    # ---- expert cache / prefetch façade (archer_prefetch_handle.cpp:195-218)
    def _pairs(self, tensor_ids):
        out, seen = [], set()
        for t in tensor_ids:
            le = self._node_of[int(t)]
            if le not in seen:
                seen.add(le)
                out.append(le)
        return out

    def replace_cache_candidates(self, tensor_ids: Sequence[int]):
        self._dispatcher.engine.replace_cache_candidates(self._pairs(tensor_ids))

    def enqueue_prefetch(self, tensor_id: int, gpu_id: int):
        l, e = self._node_of[int(tensor_id)]
        self._dispatcher.engine.enqueue_prefetch(l, e)

    def prefetch_hint(self, pairs, scores):
        self._dispatcher.engine.prefetch_hint(pairs, scores)

    def clean_up_resources(self):
        if self._dispatcher and self._dispatcher.engine:
            self._dispatcher.engine.close()


class expert_dispatcher:  # noqa: N801
    """py_archer_prefetch.cpp:84-92 / core/parallel/expert_dispatcher.h:44-82."""

    def __init__(self, num_experts: int, num_layers: int, dtype: int, expert_type: int, num_threads: int = 8,
                 handle: Optional[prefetch_handle] = None, top_k: Optional[int] = None, max_tokens: int = 4096,
                 num_slots: int = 0, **engine_kw):
        """The five positional arguments are the reference's (py_archer_prefetch.cpp:85); the keywords are optional
        extras: `handle` (default: the process's current prefetch_handle), `top_k` (default: learnt from the masks),
        workspace / cache sizing."""
        if dtype not in _INT2DTYPE:
            raise ValueError(f"dtype int {dtype} unsupported")
        self.num_experts, self.num_layers = num_experts, num_layers
        self.dtype, self.expert_type = _INT2DTYPE[dtype], expert_type
        self.num_threads = num_threads   # accepted for signature parity; the CUDA path needs no worker threads
        self.handle = handle if handle is not None else _CURRENT_HANDLE
        self.top_k = top_k
        self.max_tokens, self.num_slots, self.engine_kw = max_tokens, num_slots, engine_kw
        self.engine: Optional[MoEEngine] = None
        self._registered: Dict[Tuple[int, int], List[int]] = {}    # (layer, expert) -> tensor ids, in registration order
        self._queue: List[Tuple[int, int]] = []
        self._inputs = None
        self._expected = 0
        if self.handle is not None:
            self.handle._dispatcher = self

    def _default_k(self) -> int:
        if self.top_k:
            return self.top_k
        if self.expert_type in (L.EXPERT_SWITCH, L.EXPERT_SWITCH_GATED):
            return 1
        return 6 if self.expert_type == L.EXPERT_DEEPSEEK else 2       # grows on demand (see _fit)

    def _build_engine(self, k: int, max_tokens: int):
        """(Re)create the CUDA context for (k, max_tokens) and hand it every registered expert."""
        if self.engine is not None:
            self.engine.close()
        any_ids = next(iter(self._registered.values()))
        inter, hidden = self.handle._shape_of(int(any_ids[0]))   # first tensor is [I,H] for every supported expert type
        ratio = self.handle.device_memory_ratio if self.handle else 0.0
        self.max_tokens = max_tokens
        self.engine = MoEEngine(num_layers=self.num_layers, num_experts=self.num_experts, hidden=hidden, inter=inter,
                                top_k=k, dtype=self.dtype, expert_type=self.expert_type,
                                router=_ROUTER_OF_TYPE[self.expert_type], max_tokens=max_tokens,
                                num_slots=self.num_slots, device_memory_ratio=ratio, **self.engine_kw)
        for (l, e), ids in self._registered.items():
            self._register_in_engine(l, e, ids)

    def _register_in_engine(self, l: int, e: int, ids: List[int]):
        if self.handle.experts_on_disk:
            self.engine.register_expert_on_store(l, e, self.handle._reader_for(ids), ids)
            for t in ids:
                dict.pop(self.handle._tensors, t, None)          # the host copy is no longer needed: it can be read again
        else:
            self.engine.register_expert(l, e, [self.handle._tensors[int(t)] for t in ids])

    def _fit(self, k_needed: int, T: int):
        """The reference runs every set mask bit for any number of tokens; our context is sized by (top_k, max_tokens):
        grow it when a call needs more (rare: the first call of a model, or a longer prompt than any before)."""
        if self.engine is None or k_needed > self.engine.k or T > self.max_tokens:
            k = max(k_needed, self.engine.k if self.engine is not None else self._default_k())
            mt = self.max_tokens
            while mt < T:
                mt *= 2
            self._build_engine(k, mt)

    def register_expert(self, layer_idx: int, expert_idx: int, tensor_ids: Sequence[int]):
        """All ids belong to one expert, in named_parameters order (expert_dispatcher.cpp:160-173)."""
        if self.handle is None:
            raise RuntimeError("expert_dispatcher needs a prefetch_handle (construct it first, as model_offload.py does)")
        ids = [int(t) for t in tensor_ids]
        for t in ids:
            if t not in self.handle._tensors:
                raise RuntimeError(f"tensor id {t} was never offloaded")
            self.handle._node_of[t] = (layer_idx, expert_idx)
        self._registered[(layer_idx, expert_idx)] = ids
        if self.engine is not None:
            self._register_in_engine(layer_idx, expert_idx, ids)
        elif self.top_k:
            self._build_engine(self.top_k, self.max_tokens)    # sizes known up front: build now (and register this expert)

    def set_inputs(self, hidden_states: torch.Tensor, router_mask: torch.Tensor):
        self._inputs = (hidden_states, router_mask)   # no clones: reads are stream ordered (reference clones both)
        self._queue = []

    def set_expected_queue(self, expected_pending: int):
        self._expected = int(expected_pending)

    def enqueue_expert(self, layer_idx: int, expert_idx: int, gpu_id: int = -1, remote: bool = False):
        self._queue.append((layer_idx, expert_idx))

    def wait_expert(self):
        """[(output[n_e,H], layer, expert, hit)] for the enqueued experts; rows in ascending token order."""
        hidden, mask = self._inputs
        if not self._queue:
            return []
        layer = self._queue[0][0]
        E = self.num_experts
        x = hidden.reshape(-1, hidden.shape[-1])
        m = mask.reshape(-1, E)
        wanted = sorted({e for _, e in self._queue})
        sel = torch.zeros(E, dtype=torch.bool, device=m.device)
        sel[wanted] = True
        m = m.ne(0) & sel[None, :]
        # experts per token this call needs (the reference's own dispatch_local already synchronises here, :34-39)
        k_needed = int(m.sum(dim=-1).max().item()) if m.numel() else 0
        self._fit(max(k_needed, 1), x.shape[0])
        T = self.engine.route_from_mask(layer, x, m)
        resident_before = {e: self.engine.is_resident(layer, e) for e in wanted}
        self.engine.run_experts(layer, T)
        rows, offs = self.engine.expert_outputs(T)
        out = []
        for e in wanted:
            if offs[e + 1] > offs[e] or (layer, e) in self._queue:
                out.append((rows[offs[e]:offs[e + 1]], layer, e, int(resident_before[e])))
        self._queue = []
        return out

    def clear_expert_cache_counts(self):
        if self.engine is not None:
            self.engine.clear_expert_cache_counts()


class DistributedExpertExecutor:
    """moe_infinity/distributed/expert_executor.py:19-58 (dispatch_local only; RPC `dispatch` is dead code there)."""

    def __init__(self, archer_config=None):
        self.archer_config = archer_config
        self.expert_dispatcher = None

    def set_expert_dispatcher(self, dispatcher):
        self.expert_dispatcher = dispatcher

    def set_device_map_manager(self, device_map_manager):
        self.device_map_manager = device_map_manager

    def dispatch_local(self, hidden_states, router_mask, layer_id):
        num_expert = router_mask.shape[-1]
        expert_count = torch.sum(router_mask.view((-1, num_expert)), dim=0).cpu().numpy().flatten()
        expert_list = np.arange(num_expert).astype(int)[expert_count > 0].tolist()
        self.expert_dispatcher.set_inputs(hidden_states, router_mask)
        self.expert_dispatcher.set_expected_queue(len(expert_list))
        total_gpus = max(1, torch.cuda.device_count())
        for expert_id in expert_list:
            self.expert_dispatcher.enqueue_expert(layer_id, expert_id, expert_id % total_gpus, False)
        return self.expert_dispatcher.wait_expert()

    # fused fast path reached through the same attribute (`block.expert_executor`), SURVEY §8(b)
    def moe_forward(self, layer_id: int, hidden_states: torch.Tensor, router_logits=None, scores=None, seq_len: int = 0):
        eng = self.expert_dispatcher.engine if hasattr(self.expert_dispatcher, "engine") else self.expert_dispatcher
        return eng.forward(layer_id, hidden_states, router_logits=router_logits, scores=scores, seq_len=seq_len)

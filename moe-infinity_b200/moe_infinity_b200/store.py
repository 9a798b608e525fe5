"""The reference's on-disk tensor store, read and written in its own format ("disk tier + on-disk format").

Layout under `<prefix>/` (core/aio/archer_tensor_handle.cpp:18-19,153-156; archer_prefetch_handle.cpp:229-237):
  archer_index       binary index, `ArcherTensorIndex::Serialize` (archer_tensor_index.cpp:105-113):
                       u32 count, then per tensor:  u32 id | u32 file_id | i64 offset | u64 nbytes | i64 ndim | i64 dims[ndim] |
                       u8 pinned | u8 requires_grad | i8 c10::ScalarType | i8 device_index | i8 device_type | i8 layout
                     (entries in unordered_map order: readers must not assume any order; little endian, no padding)
  archer_param_<n>   raw tensor bytes; every tensor starts at a 4096-byte aligned offset (`kAioAlignment`,
                     archer_prio_aio_handle.h:18) and the reference reads/writes whole aligned blocks with O_DIRECT, so the
                     file always extends to the aligned end of its last tensor.  The reference only ever uses file 0.
A store written here is readable by the reference (`ArcherTensorHandle` ctor -> `Deserialize`, :41-46) and vice versa;
tests/test_store_format.py checks both directions against the reference's own (compiled) index code.

Host-side only: tensors are read into (pinned) host memory, from where `b2m_register_expert` takes over.
`ArcherTensorStore` is the format in plain Python (writer + simple reader); `NativeStoreReader` is the product's reader
(csrc/store_reader.cpp behind the C ABI: worker threads, O_DIRECT, asynchronous tickets, two priorities) and the handle
`MoEEngine.register_expert_on_store` takes for experts that stay on disk.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence

import torch

ALIGN = 4096                      # kAioAlignment
INDEX_NAME = "archer_index"       # ARCHER_IHDEX_NAME (sic)
PARAM_NAME = "archer_param"       # ARCHER_PARAM_NAME

# c10::ScalarType values (c10/core/ScalarType.h) <-> torch dtypes, for the types a checkpoint can hold
_SCALAR_TYPES = {
    0: torch.uint8, 1: torch.int8, 2: torch.int16, 3: torch.int32, 4: torch.int64, 5: torch.float16, 6: torch.float32,
    7: torch.float64, 11: torch.bool, 15: torch.bfloat16, 23: torch.float8_e5m2, 24: torch.float8_e4m3fn,
}
_SCALAR_OF = {v: k for k, v in _SCALAR_TYPES.items()}


def _align(n: int) -> int:
    return (n + ALIGN - 1) & ~(ALIGN - 1)


@dataclass
class TensorMeta:
    """TensorStorageMeta (archer_tensor_index.h:22-35), serialised fields only."""
    file_id: int
    offset: int
    size: int
    shape: List[int]
    scalar_type: int
    pinned: bool = False
    requires_grad: bool = False
    device_index: int = -1        # CPU tensors: torch::Device(kCPU).index() == -1
    device_type: int = 0          # c10::DeviceType::CPU
    layout: int = 0               # c10::Layout::Strided

    @property
    def dtype(self) -> torch.dtype:
        if self.scalar_type not in _SCALAR_TYPES:
            raise ValueError(f"c10::ScalarType {self.scalar_type} is not supported by this reader")
        return _SCALAR_TYPES[self.scalar_type]


def parse_index(data: bytes) -> Dict[int, TensorMeta]:
    """`ArcherTensorIndex::Deserialize` (archer_tensor_index.cpp:115-132)."""
    if len(data) < 4:
        raise ValueError("archer_index is truncated")
    (count,) = struct.unpack_from("<I", data, 0)
    pos = 4
    out: Dict[int, TensorMeta] = {}
    for _ in range(count):
        if pos + 32 > len(data):
            raise ValueError("archer_index is truncated")
        tid, file_id, offset, size, ndim = struct.unpack_from("<IIqQq", data, pos)
        pos += 32
        if ndim < 0 or pos + 8 * ndim + 6 > len(data):
            raise ValueError("archer_index is corrupt (bad rank)")
        shape = list(struct.unpack_from(f"<{ndim}q", data, pos))
        pos += 8 * ndim
        pinned, req, st, dev_idx, dev_type, layout = struct.unpack_from("<??bbbb", data, pos)
        pos += 6
        out[tid] = TensorMeta(file_id, offset, size, shape, st, pinned, req, dev_idx, dev_type, layout)
    return out


def serialize_index(index: Dict[int, TensorMeta]) -> bytes:
    """`ArcherTensorIndex::Serialize` (archer_tensor_index.cpp:105-113); entries in ascending id."""
    parts = [struct.pack("<I", len(index))]
    for tid in sorted(index):
        m = index[tid]
        parts.append(struct.pack("<IIqQq", tid, m.file_id, m.offset, m.size, len(m.shape)))
        parts.append(struct.pack(f"<{len(m.shape)}q", *m.shape))
        parts.append(struct.pack("<??bbbb", m.pinned, m.requires_grad, m.scalar_type, m.device_index, m.device_type, m.layout))
    return b"".join(parts)


class ArcherTensorStore:
    """`ArcherTensorHandle` restricted to its storage duties (archer_tensor_handle.cpp:23-86,153-156)."""

    def __init__(self, prefix: str):
        self.prefix = prefix if prefix.endswith("/") else prefix + "/"
        bare = self.prefix.rstrip("/") or "/"
        if os.path.exists(bare) and not os.path.isdir(bare):
            raise ValueError(f"Invalid prefix: {self.prefix} is not a directory")      # :32-34
        os.makedirs(self.prefix, exist_ok=True)                                        # :35-38
        self.index: Dict[int, TensorMeta] = {}
        self._initialized = False
        self._file_id = 0
        self._file_offset = 0
        self._dirty = False
        path = self.index_path
        if os.path.exists(path):                                                        # :41-46
            with open(path, "rb") as f:
                self.index = parse_index(f.read())
            self._initialized = True
            # the reference restarts at offset 0 after a reload (file_offset_ is not persisted) and would overwrite the
            # first tensors when new ids are added to an existing store; appending after the last stored byte instead
            for m in self.index.values():
                if m.file_id == self._file_id:
                    self._file_offset = max(self._file_offset, m.offset + _align(m.size))

    # ---- names
    @property
    def index_path(self) -> str:
        return self.prefix + INDEX_NAME

    def param_path(self, file_id: int) -> str:                                         # GetIndexFileName :153-156
        return f"{self.prefix}{PARAM_NAME}_{file_id}"

    # ---- queries (is_tensor_index_initialized / is_tensor_offloaded, py_archer_prefetch.cpp:60-66)
    def is_initialized(self) -> bool:
        return self._initialized

    def __contains__(self, tensor_id: int) -> bool:
        return int(tensor_id) in self.index

    def __len__(self) -> int:
        return len(self.index)

    def aligned_size(self, tensor_id: int) -> int:                                     # GetTensorSizeAligned :88-98
        return _align(self.index[int(tensor_id)].size)

    # ---- writing (StoreTensor :53-86 + OffloadTensor's Serialize, archer_prefetch_handle.cpp:229-237)
    def store_tensor(self, tensor_id: int, tensor: torch.Tensor, flush: bool = True):
        tensor_id = int(tensor_id)
        t = tensor.detach().to("cpu").contiguous()
        if t.dtype not in _SCALAR_OF:
            raise ValueError(f"dtype {t.dtype} cannot be stored")
        nbytes = t.numel() * t.element_size()
        old = self.index.get(tensor_id)
        if old is not None:
            if old.size != nbytes:                                                      # :70-74 (the reference aborts)
                raise ValueError(f"Tensor {tensor_id} size mismatch {old.size} != {nbytes}")
            meta = old
        else:
            meta = TensorMeta(self._file_id, self._file_offset, nbytes, list(t.shape), _SCALAR_OF[t.dtype],
                              pinned=t.is_pinned(), requires_grad=bool(tensor.requires_grad))
            self._file_offset += _align(nbytes)
            self.index[tensor_id] = meta
        raw = t.reshape(-1).view(torch.uint8).numpy().tobytes() if nbytes else b""
        fd = os.open(self.param_path(meta.file_id), os.O_RDWR | os.O_CREAT, 0o644)
        try:
            os.pwrite(fd, raw, meta.offset)
            end = meta.offset + _align(nbytes)
            if os.fstat(fd).st_size < end:
                os.ftruncate(fd, end)             # whole aligned blocks, as the O_DIRECT writer leaves them
        finally:
            os.close(fd)
        self._dirty = True
        if flush:
            self.flush()

    def flush(self):
        """Write archer_index (the reference re-serialises it after EVERY offload call -- O(n^2) in the tensor count,
        SURVEY §8 c.2; pass flush=False to store_tensor and call this once)."""
        if not self._dirty and os.path.exists(self.index_path):
            return
        tmp = self.index_path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(serialize_index(self.index))
        os.replace(tmp, self.index_path)
        self._dirty = False
        self._initialized = True

    # ---- reading
    def read_into(self, tensor_id: int, out: torch.Tensor) -> torch.Tensor:
        """Read the tensor's bytes into `out` (a CPU uint8 tensor of at least `size` bytes, e.g. a slice of a pinned blob)."""
        m = self.index[int(tensor_id)]
        buf = out.reshape(-1).view(torch.uint8)
        if buf.numel() < m.size:
            raise ValueError(f"buffer of {buf.numel()} bytes is too small for tensor {tensor_id} ({m.size} bytes)")
        if m.size:
            with open(self.param_path(m.file_id), "rb", buffering=0) as f:
                f.seek(m.offset)
                mv = memoryview(buf.numpy())[: m.size]
                got = 0
                while got < m.size:
                    n = f.readinto(mv[got:])
                    if not n:
                        raise IOError(f"{self.param_path(m.file_id)} is truncated (tensor {tensor_id})")
                    got += n
        return out

    def read_tensor(self, tensor_id: int, pin: bool = False) -> torch.Tensor:
        m = self.index[int(tensor_id)]
        raw = torch.empty(m.size, dtype=torch.uint8, pin_memory=pin)
        self.read_into(tensor_id, raw)
        return raw.view(m.dtype).reshape(m.shape)

    def read_expert_blob(self, tensor_ids: Sequence[int], out: Optional[torch.Tensor] = None, pin: bool = False) -> torch.Tensor:
        """The expert's tensors concatenated in `tensor_ids` order without padding -- the blob `b2m_register_expert` takes
        (the reference's own host blob pads every tensor to 4096 bytes, model_topology.cpp:429-431)."""
        total = sum(self.index[int(t)].size for t in tensor_ids)
        if out is None:
            out = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
        elif out.numel() * out.element_size() < total:
            raise ValueError("blob buffer too small")
        flat = out.reshape(-1).view(torch.uint8)
        off = 0
        for t in tensor_ids:
            n = self.index[int(t)].size
            self.read_into(t, flat[off:off + n])
            off += n
        return out

    def ids(self) -> Iterable[int]:
        return sorted(self.index)


class NativeStoreReader:
    """`b2m_store_*` (include/b2m.h): the disk tier's reader.  What `ArcherTensorHandle::ReadTensor` ->
    `ArcherPrioAioHandle::Read` is in the reference (archer_tensor_handle.cpp:189-201, archer_prio_aio_handle.cpp:37-70),
    asynchronous and multi-threaded.  Needs libb2m.so but no GPU."""

    def __init__(self, prefix: str, num_threads: int = 0, block_bytes: int = 0, odirect: bool = True):
        import ctypes as C
        from . import _lib as L
        self._C, self._L, self._lib = C, L, L.load()
        self._h = C.c_void_p()
        rc = self._lib.b2m_store_open(os.fsencode(prefix.rstrip("/") or "/"), num_threads, block_bytes,
                                      0 if odirect else L.STORE_NO_ODIRECT, C.byref(self._h))
        if rc:
            self._h = C.c_void_p()
            raise IOError(f"b2m_store_open({prefix!r}) failed with {rc}"
                          + (": no readable archer_index" if rc == L.B2M_EIO else ": corrupt archer_index"))

    @property
    def handle(self):
        return self._h

    def _ck(self, rc):
        if rc < 0:
            msg = self._lib.b2m_store_last_error(self._h)
            raise (IOError if rc == self._L.B2M_EIO else ValueError)(f"store error {rc}: {msg.decode() if msg else ''}")
        return rc

    def _ids(self, tensor_ids):
        arr = (self._C.c_uint32 * len(tensor_ids))(*[int(t) for t in tensor_ids])
        return arr, len(tensor_ids)

    def __len__(self):
        return self._ck(self._lib.b2m_store_count(self._h))

    def tensor(self, tensor_id: int):
        """(file_id, offset, nbytes) of a tensor."""
        C = self._C
        f, o, n = C.c_uint32(), C.c_int64(), C.c_uint64()
        self._ck(self._lib.b2m_store_tensor(self._h, int(tensor_id), C.byref(f), C.byref(o), C.byref(n)))
        return f.value, o.value, n.value

    def blob_bytes(self, tensor_ids: Sequence[int]) -> int:
        arr, n = self._ids(tensor_ids)
        tot = self._C.c_uint64()
        self._ck(self._lib.b2m_store_blob_bytes(self._h, arr, n, self._C.byref(tot)))
        return tot.value

    def read_async(self, tensor_ids: Sequence[int], out: torch.Tensor, high_priority: bool = False,
                   blob_offset: int = 0, nbytes: Optional[int] = None) -> int:
        """Start reading the concatenation of `tensor_ids` (or bytes [blob_offset, +nbytes) of it) into the CPU tensor `out`
        (keep it alive until `wait`); returns a ticket."""
        if out.device.type != "cpu" or not out.is_contiguous():
            raise ValueError("the destination must be a contiguous CPU tensor")
        arr, n = self._ids(tensor_ids)
        cap = out.numel() * out.element_size()
        tk = self._C.c_uint64()
        if blob_offset == 0 and nbytes is None:
            self._ck(self._lib.b2m_store_read_async(self._h, arr, n, out.data_ptr(), cap, int(high_priority), self._C.byref(tk)))
        else:
            nbytes = self.blob_bytes(tensor_ids) - blob_offset if nbytes is None else nbytes
            if nbytes > cap:
                raise ValueError("destination too small")
            self._ck(self._lib.b2m_store_read_range_async(self._h, arr, n, blob_offset, nbytes, out.data_ptr(),
                                                         int(high_priority), self._C.byref(tk)))
        return tk.value

    def poll(self, ticket: int) -> bool:
        return self._ck(self._lib.b2m_store_poll(self._h, ticket)) == 1

    def wait(self, ticket: int) -> None:
        self._ck(self._lib.b2m_store_wait(self._h, ticket))

    def read_expert_blob(self, tensor_ids: Sequence[int], out: Optional[torch.Tensor] = None, pin: bool = False) -> torch.Tensor:
        """Same result as `ArcherTensorStore.read_expert_blob`, through the native reader."""
        total = self.blob_bytes(tensor_ids)
        if out is None:
            out = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
        self.wait(self.read_async(tensor_ids, out, high_priority=True))
        return out

    def stats(self) -> Dict[str, int]:
        out = (self._C.c_uint64 * 4)()
        self._ck(self._lib.b2m_store_stats(self._h, out))
        return dict(bytes_read=out[0], direct_blocks=out[1], buffered_blocks=out[2], requests=out[3])

    def close(self):
        if self._h:
            self._lib.b2m_store_close(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

"""Disk tier reader (csrc/store_reader.cpp behind include/b2m.h's b2m_store_*) on the CPU: the product's
library is loaded as it is -- the reader is host code and needs no GPU -- and checked against the plain-Python statement of
the reference's store format (moe_infinity_b200/store.py, itself pinned on the reference's compiled index code by
tests/test_store_format.py).  Reference behaviour being replaced: ArcherTensorHandle::ReadTensor
(core/aio/archer_tensor_handle.cpp:189-201) -> ArcherPrioAioHandle::Read (archer_prio_aio_handle.cpp:37-70)."""
import os
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moe-infinity_b200"))

from moe_infinity_b200 import _lib as L  # noqa: E402
from moe_infinity_b200.store import ArcherTensorStore, NativeStoreReader  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.exists(L.LIB_PATH), reason="libb2m.so is not built")


def _make_store(path, sizes, seed=0):
    """Tensors of the given byte sizes (uint8), ids 10, 11, ...; returns {id: bytes}."""
    st = ArcherTensorStore(str(path))
    g = torch.Generator().manual_seed(seed)
    data = {}
    for i, n in enumerate(sizes):
        t = torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g)
        st.store_tensor(10 + i, t, flush=False)
        data[10 + i] = t.numpy().tobytes()
    st.flush()
    return data


@pytest.mark.parametrize("odirect", [True, False])
@pytest.mark.parametrize("block", [4096, 1 << 16, 0])
def test_reads_match_the_python_store(tmp_path, odirect, block):
    sizes = [4096 * 40, 5, 4096 * 3 + 17, 0, 1 << 20, 4095, 4097]        # aligned, tiny, ragged, empty, large
    data = _make_store(tmp_path, sizes)
    rd = NativeStoreReader(str(tmp_path), num_threads=4, block_bytes=block, odirect=odirect)
    assert len(rd) == len(sizes)
    py = ArcherTensorStore(str(tmp_path))
    for tid, raw in data.items():
        f, off, n = rd.tensor(tid)
        assert (f, off, n) == (py.index[tid].file_id, py.index[tid].offset, len(raw))
        assert off % 4096 == 0
    # one tensor at a time, then the "expert blob" = several tensors back to back in a caller-chosen order
    for tid, raw in data.items():
        out = torch.full((len(raw) + 8,), 0xAB, dtype=torch.uint8)
        rd.wait(rd.read_async([tid], out))
        assert out[: len(raw)].numpy().tobytes() == raw
        assert (out[len(raw):] == 0xAB).all()                              # never writes past the tensor's end
    order = [14, 10, 12, 11, 16, 13, 15]
    blob = rd.read_expert_blob(order)
    assert blob.numpy().tobytes() == b"".join(data[t] for t in order)
    assert blob.numpy().tobytes() == py.read_expert_blob(order).numpy().tobytes()
    st = rd.stats()
    assert st["bytes_read"] == sum(sizes) + sum(len(data[t]) for t in order)
    if not odirect:
        assert st["direct_blocks"] == 0
    rd.close()


def test_byte_ranges_of_a_blob_cross_tensor_boundaries(tmp_path):
    sizes = [4096 * 5, 4096 * 2 + 100, 4096 * 7, 333]
    data = _make_store(tmp_path, sizes, seed=1)
    ids = [10, 11, 12, 13]
    whole = b"".join(data[t] for t in ids)
    rd = NativeStoreReader(str(tmp_path), num_threads=3, block_bytes=8192)
    assert rd.blob_bytes(ids) == len(whole)
    rng = np.random.default_rng(0)
    cuts = [(0, len(whole)), (0, 0), (len(whole), 0), (4096 * 5 - 1, 2), (4096, 4096 * 9)]             # (offset, length)
    for _ in range(40):
        a, b = sorted(int(v) for v in rng.integers(0, len(whole) + 1, 2))
        cuts.append((a, b - a))
    for a, n in cuts:
        out = torch.zeros(max(n, 1), dtype=torch.uint8)
        rd.wait(rd.read_async(ids, out, blob_offset=a, nbytes=n))
        assert out[:n].numpy().tobytes() == whole[a:a + n], (a, n)
    # a staging-chunk walk like api.cu's: fixed chunks into a small ring, two reads in flight
    chunk = 4096 * 3
    ring = [torch.zeros(chunk, dtype=torch.uint8) for _ in range(3)]
    got = bytearray()
    nch = (len(whole) + chunk - 1) // chunk
    tk = {}
    for ci in range(min(2, nch)):
        tk[ci] = rd.read_async(ids, ring[ci % 3], True, ci * chunk, min(chunk, len(whole) - ci * chunk))
    for ci in range(nch):
        rd.wait(tk.pop(ci))
        got += ring[ci % 3][: min(chunk, len(whole) - ci * chunk)].numpy().tobytes()
        if ci + 2 < nch:
            tk[ci + 2] = rd.read_async(ids, ring[(ci + 2) % 3], True, (ci + 2) * chunk, min(chunk, len(whole) - (ci + 2) * chunk))
    assert bytes(got) == whole
    rd.close()


def test_errors_are_returned_not_fatal(tmp_path):
    data = _make_store(tmp_path, [4096 * 4, 100])
    with pytest.raises(IOError):
        NativeStoreReader(str(tmp_path / "nowhere"))
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "archer_index").write_bytes(b"\x05\x00\x00\x00garbage")
    with pytest.raises(IOError):
        NativeStoreReader(str(bad))
    rd = NativeStoreReader(str(tmp_path), num_threads=2)
    out = torch.zeros(4096 * 4, dtype=torch.uint8)
    with pytest.raises(ValueError):
        rd.read_async([99], out)                                            # unknown tensor id
    with pytest.raises(ValueError):
        rd.read_async([10, 11], out)                                        # destination too small
    with pytest.raises(ValueError):
        rd.read_async([10], out, blob_offset=4096 * 4, nbytes=1)            # range past the blob's end
    with pytest.raises(ValueError):
        rd.wait(12345)                                                      # unknown ticket
    t = rd.read_async([10], out)
    rd.wait(t)
    with pytest.raises(ValueError):
        rd.wait(t)                                                          # a ticket is retired by its wait
    # the data file is shorter than the index says: B2M_EIO from wait(), and the reader stays usable
    with open(tmp_path / "archer_param_0", "r+b") as f:
        f.truncate(4096 * 2)
    t = rd.read_async([10], out)
    with pytest.raises(IOError):
        rd.wait(t)
    half = torch.zeros(4096 * 2, dtype=torch.uint8)
    rd.wait(rd.read_async([10], half, blob_offset=0, nbytes=4096 * 2))
    assert half.numpy().tobytes() == data[10][: 4096 * 2]
    rd.close()
    rd.close()                                                              # idempotent


def test_many_requests_from_many_threads_and_both_priorities(tmp_path):
    sizes = [4096 * 16 + (i * 37) % 4096 for i in range(24)]
    data = _make_store(tmp_path, sizes, seed=3)
    rd = NativeStoreReader(str(tmp_path), num_threads=6, block_bytes=4096 * 2)
    errs = []

    def client(k):
        try:
            rng = np.random.default_rng(k)
            for it in range(30):
                ids = [int(x) for x in rng.choice(np.arange(10, 10 + len(sizes)), size=int(rng.integers(1, 4)), replace=False)]
                want = b"".join(data[t] for t in ids)
                out = torch.zeros(len(want), dtype=torch.uint8)
                tk = rd.read_async(ids, out, high_priority=bool(it % 2))
                while not rd.poll(tk) and it % 3 == 0:
                    pass
                rd.wait(tk)
                if out.numpy().tobytes() != want:
                    errs.append((k, it, ids))
        except Exception as ex:  # pragma: no cover
            errs.append((k, repr(ex)))

    ths = [threading.Thread(target=client, args=(k,)) for k in range(6)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs[:3]
    assert rd.stats()["requests"] == 6 * 30
    # close() with requests still queued serves them first (no buffer is left half written by a dying worker)
    outs = [torch.zeros(sizes[i], dtype=torch.uint8) for i in range(len(sizes))]
    for i in range(len(sizes)):
        rd.read_async([10 + i], outs[i], high_priority=False)
    rd.close()
    for i in range(len(sizes)):
        assert outs[i].numpy().tobytes() == data[10 + i]


def test_aligned_destinations_take_o_direct_where_the_file_system_has_it(tmp_path):
    n = 1 << 22
    data = _make_store(tmp_path, [n, 4096 * 3 + 5], seed=9)
    rd = NativeStoreReader(str(tmp_path), num_threads=4, block_bytes=1 << 20)
    raw = torch.zeros(n + 4096, dtype=torch.uint8)
    off = (-raw.data_ptr()) % 4096
    out = raw[off:off + n]                                                  # 4096-aligned, like a pinned staging chunk
    rd.wait(rd.read_async([10], out))
    assert out.numpy().tobytes() == data[10]
    st = rd.stats()
    assert st["direct_blocks"] + st["buffered_blocks"] == 4                 # 4 MiB in 1 MiB blocks, whichever descriptor served them
    tail = torch.zeros(4096 * 4, dtype=torch.uint8)
    rd.wait(rd.read_async([11], tail))                                      # ragged size: its last block is always buffered
    assert tail[: 4096 * 3 + 5].numpy().tobytes() == data[11]
    assert rd.stats()["buffered_blocks"] >= 1
    rd.close()

"""Throughput of the disk tier's native reader (csrc/store_reader.cpp) on this machine's disk, next to the reference's
configuration of the same job (one worker, 1 MiB blocks: core/aio/archer_aio_utils.cpp:13, archer_prio_aio_handle.cpp:113-121)
and the plain-Python reader.  CPU only.  `python tools/bench_store.py [--dir D] [--mb 1024]` prints one JSON line."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moe-infinity_b200"))
from moe_infinity_b200.store import ArcherTensorStore, NativeStoreReader  # noqa: E402


def aligned(nbytes):
    raw = torch.empty(nbytes + 4096, dtype=torch.uint8)
    off = (-raw.data_ptr()) % 4096
    return raw[off:off + nbytes]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=None)
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--tensor-mb", type=int, default=112)        # a Mixtral expert matrix is 112 MiB
    a = ap.parse_args()
    d = tempfile.mkdtemp(dir=a.dir)
    try:
        st = ArcherTensorStore(d)
        n = max(1, a.mb // a.tensor_mb)
        tb = a.tensor_mb << 20
        src = torch.randint(0, 256, (tb,), dtype=torch.uint8)
        for i in range(n):
            st.store_tensor(i, src.roll(i), flush=False)
        st.flush()
        os.sync()
        ids = list(range(n))
        out = aligned(n * tb)
        res = {}
        for name, kw in (("native 8 threads x 4 MiB O_DIRECT", dict(num_threads=8, block_bytes=4 << 20)),
                         ("native 16 threads x 4 MiB O_DIRECT", dict(num_threads=16, block_bytes=4 << 20)),
                         ("native 8 threads x 4 MiB buffered", dict(num_threads=8, block_bytes=4 << 20, odirect=False)),
                         ("reference shape: 1 thread x 1 MiB O_DIRECT", dict(num_threads=1, block_bytes=1 << 20))):
            rd = NativeStoreReader(d, **kw)
            best = 0.0
            for _ in range(3):
                t0 = time.perf_counter()
                rd.wait(rd.read_async(ids, out, high_priority=True))
                best = max(best, n * tb / (time.perf_counter() - t0) / 1e9)
            s = rd.stats()
            res[name] = {"GB/s": round(best, 2), "direct_blocks": s["direct_blocks"], "buffered_blocks": s["buffered_blocks"]}
            rd.close()
        assert out[:tb].equal(src) and out[(n - 1) * tb:].equal(src.roll(n - 1))
        t0 = time.perf_counter()
        st.read_expert_blob(ids, out)
        res["python reader (store.py)"] = {"GB/s": round(n * tb / (time.perf_counter() - t0) / 1e9, 2)}
        print(json.dumps({"bytes": n * tb, "dir": d, "cpu_count": os.cpu_count(), "results": res}))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()

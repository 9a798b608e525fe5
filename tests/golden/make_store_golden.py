"""Generate tests/golden/archer_index_ref.bin with the reference's OWN index writer (dev container only).

Run:  python tests/golden/make_store_golden.py
`ArcherTensorIndex::Serialize` (core/aio/archer_tensor_index.cpp:105-113, compiled as-is into oracle/_ref) over metas
built the way `ArcherTensorHandle::StoreTensor` builds them (archer_tensor_handle.cpp:53-86).  The entries are listed in
STORE_GOLDEN_ENTRIES so tests/test_store_format.py can check the parser without the compiled reference."""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_module  # noqa: E402

# (tensor id, file id, offset, shape, dtype)
STORE_GOLDEN_ENTRIES = [
    (0, 0, 0, [32000, 64], torch.bfloat16),
    (1, 0, 4096000, [64], torch.float32),
    (7, 0, 4100096, [14336, 4096], torch.bfloat16),
    (8, 0, 121540608, [4096, 14336], torch.float16),
    (4000000000, 0, 238981120, [], torch.int64),
    (12, 0, 238985216, [3, 5, 7], torch.float8_e4m3fn),
    (13, 0, 238989312, [0], torch.uint8),
]


def main():
    R = ref_module.load()
    assert R is not None, "oracle/_ref/ref_expert_module.so missing (python __graft_entry__.py builds it)"
    # only metadata is serialised (nbytes, sizes, options): big shapes are expanded views of one element, no memory
    entries = [(i, f, off, torch.zeros(1, dtype=dt).expand(shape) if len(shape) and max(shape) >= 100 else torch.zeros(shape, dtype=dt))
               for i, f, off, shape, dt in STORE_GOLDEN_ENTRIES]
    R.index_serialize(os.path.join(HERE, "archer_index_ref.bin"), entries)
    print(sorted(R.index_deserialize(os.path.join(HERE, "archer_index_ref.bin"))))


if __name__ == "__main__":
    main()

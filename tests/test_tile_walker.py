"""CPU: the persistent grouped GEMM's tile scheduling (csrc/tile_walker.cuh, shared with the kernel) replayed on the host.

tests/host/tile_walker_check.cpp walks every CTA exactly as grouped_gemm.cu does -- tile table from the routing offsets,
strided tile ids, the stream-K unit ranges of the split-K down projection, the 2-CTA cluster variant -- over seeded random
routing tables (empty experts, non-resident experts, ragged token counts, 16..256-token tiles) and demands that every
(expert, weight-row tile, token tile, k-block) unit is processed exactly once."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_every_unit_is_processed_exactly_once(tmp_path):
    exe = tmp_path / "tile_walker_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "moe-infinity_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "tile_walker_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), "4000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr

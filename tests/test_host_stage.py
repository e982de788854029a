"""Host-side staging pool (boxmot_b200/csrc/host_stage.h): the parallel copy of pageable frames into the page-locked
buffer is exercised by a standalone C++ driver -- 2 000 randomised sizes / alignments against memcmp, in-order
`landed` callbacks, an exception thrown from a callback, re-use afterwards."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_stage_pool_copies_are_exact(tmp_path):
    exe = tmp_path / "stage_pool_test"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "stage_pool_test.cpp", "-o", str(exe)],
                          cwd=ROOT / "scripts" / "microbench")
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")

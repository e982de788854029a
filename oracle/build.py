"""Compile the oracle's C restatements with gcc (checker build; not part of the product)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"


def build_oracle(force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    src = HERE / "lapjv.c"
    out = BUILD / "liboracle_lapjv.so"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        tmp = BUILD / f".liboracle_lapjv.{os.getpid()}.so"
        subprocess.check_call(["gcc", "-O2", "-Wall", "-shared", "-fPIC", "-o", str(tmp), str(src)])
        os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build_oracle(force=True))

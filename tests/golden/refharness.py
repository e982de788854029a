"""Import the UNMODIFIED reference (/root/reference) in this container to generate golden vectors.

Only used by tests/golden/make_golden.py and by optional `-m "not gpu"` cross-checks that skip when
/root/reference is absent (it does not exist on the GPU box).  Missing third-party modules are replaced by
inert stubs (`gdown`, `ftfy`, `yacs`: non-arithmetic) and `lap` by the oracle restatement (oracle/lap.py) --
SURVEY.md section 8(c).
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("BOXMOT_REFERENCE_ROOT", "/root/reference"))
REPO_ROOT = Path(__file__).resolve().parents[2]


def reference_available() -> bool:
    return (REFERENCE_ROOT / "boxmot" / "trackers" / "basetracker.py").is_file()


def install_reference() -> None:
    """Make `import boxmot` resolve to the reference, with stubs for absent third-party modules."""
    if "boxmot" in sys.modules and getattr(sys.modules["boxmot"], "__file__", "").startswith(str(REFERENCE_ROOT)):
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    os.environ.pop("GITHUB_ACTIONS", None)
    if str(REPO_ROOT) not in sys.path:
        sys.path.insert(0, str(REPO_ROOT))
    from oracle import lap as oracle_lap

    lap_mod = types.ModuleType("lap")
    lap_mod.lapjv = oracle_lap.lapjv
    lap_mod.__version__ = "0.9.4-oracle-restatement"
    sys.modules.setdefault("lap", lap_mod)
    for name in ("gdown", "ftfy"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        yacs_config = types.ModuleType("yacs.config")

        class CfgNode(dict):
            def __getattr__(self, k):
                return self[k]

            def __setattr__(self, k, v):
                self[k] = v

        yacs_config.CfgNode = CfgNode
        yacs.config = yacs_config
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = yacs_config
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))

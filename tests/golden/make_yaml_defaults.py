"""Dump the `default:` value of every parameter of the reference's tracker YAMLs
(`boxmot/configs/trackers/{bytetrack,botsort,deepocsort,strongsort}.yaml`, read by `create_tracker` through
`get_tracker_config`) to tests/golden/tracker_yaml_defaults.json.   Run: python tests/golden/make_yaml_defaults.py"""
import json
from pathlib import Path

import yaml

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/boxmot/configs/trackers")



def walk(node, acc):
    """Every `<param>: {default: ...}` entry, nested conditional parameters included (what the reference's
    `flatten_yaml_config` hands to the tracker constructor)."""
    if isinstance(node, dict):
        for k, v in node.items():
            if isinstance(v, dict) and "default" in v:
                acc[str(k)] = v["default"]
            walk(v, acc)
    elif isinstance(node, list):
        for v in node:
            walk(v, acc)
    return acc


out = {}
for name in ("bytetrack", "botsort", "deepocsort", "strongsort"):
    out[name] = walk(yaml.safe_load((REF / f"{name}.yaml").read_text()), {})
(HERE / "tracker_yaml_defaults.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
print({k: len(v) for k, v in out.items()})

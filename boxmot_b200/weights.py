"""ReID weight preparation: reference state dict (.pt) -> BN-folded `.b200reid` blob for the CUDA kernels.

Plays the role of the reference's automatic `.pt -> .onnx` export for its native ReID path
(boxmot/native/_common.py:453-570) and of its checkpoint loader (boxmot/reid/core/registry.py:126-164:
unwrap `state_dict`, strip `module.`).  Pure tensor plumbing on the host; no inference happens here.

Blob layout (little endian): 16 int32 header words
    [magic 'B2RE', version, arch (1 = OSNet), c0, c1, c2, c3, feat_dim, n_floats, 0...]
followed by float32 parameters in the exact order csrc/reid_model.cu walks them:
    stem      W[147][c0] (k = (kh*7+kw)*3 + ci, BN folded), b[c0]
    per stage s=0..2, per block j=0..1 (cin, cout, mid = cout/4, hid = mid/16):
        conv1      W[cin][mid], b[mid]
        10 x light (a0 | b0 b1 | c0 c1 c2 | d0 d1 d2 d3):  pw W[mid][mid], dw W[9][mid] (BN folded), b[mid]
        gate       fc1 W[mid][hid], b[hid], fc2 W[hid][mid], b[mid]
        combine    W[mid (+ cin if downsample)][cout]  (conv3 rows, then downsample rows), b[cout] (summed)
      transition (s < 2)  W[cout][cout], b[cout]
    conv5     W[c3][c3], b[c3]
    fc        W[c3][feat] (BatchNorm1d folded), b[feat]
All 1x1 weights are stored K-major ([cin][cout]) so a thread owning consecutive output channels loads
consecutive floats; every tensor is zero-padded to a multiple of 4 floats (16-byte aligned float4 loads).
"""
from __future__ import annotations

import hashlib
import struct
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

MAGIC = 0x45523242  # 'B2RE'
VERSION = 1
ARCH_OSNET = 1
ARCH_MOBILENETV2 = 2
BRANCHES = (("conv2a", 1), ("conv2b", 2), ("conv2c", 3), ("conv2d", 4))
EPS = 1e-5


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float64)


def load_state_dict(path) -> Dict[str, np.ndarray]:
    import torch

    ckpt = torch.load(str(path), map_location="cpu", weights_only=False)
    sd = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def _bn_fold(sd, name) -> Tuple[np.ndarray, np.ndarray]:
    scale = _np(sd[name + ".weight"]) / np.sqrt(_np(sd[name + ".running_var"]) + EPS)
    shift = _np(sd[name + ".bias"]) - _np(sd[name + ".running_mean"]) * scale
    return scale, shift


def _pw(sd, name, bn=None):
    """1x1 conv weight [co][ci][1][1] -> K-major [ci][co] with an optional BN folded in; returns (W, b)."""
    w = _np(sd[name + ".weight"])[:, :, 0, 0]
    co = w.shape[0]
    b = np.zeros(co)
    if (name + ".bias") in sd:
        b = _np(sd[name + ".bias"])
    if bn is not None:
        scale, shift = _bn_fold(sd, bn)
        w = w * scale[:, None]
        b = b * scale + shift
    return w.T.copy(), b


def fold_osnet(sd) -> Tuple[List[int], List[np.ndarray]]:
    c0 = sd["conv1.conv.weight"].shape[0]
    chans = [c0] + [sd[f"conv{s + 2}.1.conv3.conv.weight"].shape[0] for s in range(3)]
    feat = sd["fc.0.weight"].shape[0]
    out: List[np.ndarray] = []
    # stem
    w = _np(sd["conv1.conv.weight"])  # [c0][3][7][7]
    scale, shift = _bn_fold(sd, "conv1.bn")
    w = w * scale[:, None, None, None]
    out += [w.transpose(2, 3, 1, 0).reshape(147, c0), shift]
    for s in range(3):
        stage = f"conv{s + 2}"
        for j in range(2):
            name = f"{stage}.{j}"
            cin = chans[s] if j == 0 else chans[s + 1]
            cout = chans[s + 1]
            mid = cout // 4
            assert sd[name + ".conv1.conv.weight"].shape[:2] == (mid, cin)
            out += list(_pw(sd, name + ".conv1.conv", name + ".conv1.bn"))
            for br, depth in BRANCHES:
                for k in range(depth):
                    lname = f"{name}.{br}" if br == "conv2a" else f"{name}.{br}.{k}"
                    wpw, _ = _pw(sd, lname + ".conv1")
                    sc, sh = _bn_fold(sd, lname + ".bn")
                    wdw = _np(sd[lname + ".conv2.weight"])[:, 0] * sc[:, None, None]  # [c][3][3]
                    out += [wpw, wdw.reshape(mid, 9).T.copy(), sh]
            w1, b1 = _pw(sd, name + ".gate.fc1")
            w2, b2 = _pw(sd, name + ".gate.fc2")
            out += [w1, b1, w2, b2]
            w3, b3 = _pw(sd, name + ".conv3.conv", name + ".conv3.bn")
            if (name + ".downsample.conv.weight") in sd:
                wd, bd = _pw(sd, name + ".downsample.conv", name + ".downsample.bn")
                out += [np.concatenate([w3, wd], 0), b3 + bd]
            else:
                assert cin == cout
                out += [w3, b3]
        if s < 2:
            out += list(_pw(sd, f"{stage}.2.0.conv", f"{stage}.2.0.bn"))
    out += list(_pw(sd, "conv5.conv", "conv5.bn"))
    wf = _np(sd["fc.0.weight"])  # [feat][c3]
    bf = _np(sd["fc.0.bias"])
    scale, shift = _bn_fold(sd, "fc.1")
    out += [(wf * scale[:, None]).T.copy(), bf * scale + shift]
    return chans + [feat], out


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _pad_mat(w: np.ndarray, rows: int, cols: int) -> np.ndarray:
    out = np.zeros((rows, cols))
    out[: w.shape[0], : w.shape[1]] = w
    return out


def _pad_vec(b: np.ndarray, n: int) -> np.ndarray:
    out = np.zeros(n)
    out[: b.shape[0]] = b
    return out


def fold_mobilenetv2(sd):
    """MobileNetV2 (reid/backbones/mobilenetv2.py) -> (stem_c, feat, block table, arrays).  Channel counts such as
    22, 33, 89, 134 are zero-padded to multiples of 4 (padded lanes stay exactly 0 through ReLU6 and feed zero
    weight rows), so the float4 kernels apply unchanged.  Blob (arch 2): header, int32 table [n_blocks][4] =
    (cin, cout, t, stride), then  stem W[27][C0p], b | per block: expand W[cinp][midp], b; dw W[9][midp], b;
    project W[midp][coutp], b | conv9 W[clastp][featp], b."""
    from .synthetic import MOBILENETV2_LAYERS

    stem_c = sd["conv1.conv.weight"].shape[0]
    arrays, table = [], []
    w = _np(sd["conv1.conv.weight"])  # [c0][3][3][3]
    sc, sh = _bn_fold(sd, "conv1.bn")
    w = (w * sc[:, None, None, None]).transpose(2, 3, 1, 0).reshape(27, stem_c)
    arrays += [_pad_mat(w, 27, _pad4(stem_c)), _pad_vec(sh, _pad4(stem_c))]
    stage = 2
    while f"conv{stage}.0.conv1.conv.weight" in sd:
        i = 0
        while f"conv{stage}.{i}.conv1.conv.weight" in sd:
            name = f"conv{stage}.{i}"
            mid, cin = sd[name + ".conv1.conv.weight"].shape[:2]
            cout = sd[name + ".conv3.0.weight"].shape[0]
            stride = MOBILENETV2_LAYERS[stage - 2][3] if i == 0 else 1
            table.append((cin, cout, mid // cin, stride))
            we, be = _pw(sd, name + ".conv1.conv", name + ".conv1.bn")
            arrays += [_pad_mat(we, _pad4(cin), _pad4(mid)), _pad_vec(be, _pad4(mid))]
            sc, sh = _bn_fold(sd, name + ".dwconv2.bn")
            wd = (_np(sd[name + ".dwconv2.conv.weight"])[:, 0] * sc[:, None, None]).reshape(mid, 9).T
            arrays += [_pad_mat(wd, 9, _pad4(mid)), _pad_vec(sh, _pad4(mid))]
            # conv3 is nn.Sequential(Conv2d, BatchNorm2d): keys conv3.0 / conv3.1
            w3 = _np(sd[name + ".conv3.0.weight"])[:, :, 0, 0]
            sc, sh = _bn_fold(sd, name + ".conv3.1")
            arrays += [_pad_mat((w3 * sc[:, None]).T, _pad4(mid), _pad4(cout)), _pad_vec(sh, _pad4(cout))]
            i += 1
        stage += 1
    w9, b9 = _pw(sd, "conv9.conv", "conv9.bn")
    feat = w9.shape[1]
    arrays += [_pad_mat(w9, _pad4(w9.shape[0]), _pad4(feat)), _pad_vec(b9, _pad4(feat))]
    return stem_c, feat, table, arrays


def export_blob(weights, out_path=None) -> Path:
    """`weights`: path to a .pt checkpoint or an in-memory state dict.  Returns the blob path."""
    if isinstance(weights, (str, Path)):
        src = Path(weights)
        if src.suffix == ".b200reid":
            return src
        sd = load_state_dict(src)
        if out_path is None:
            out_path = src.with_suffix(".b200reid")
    else:
        sd = weights
        if out_path is None:
            raise ValueError("out_path is required when exporting an in-memory state dict")
    table = []
    if "conv9.conv.weight" in sd:
        stem_c, feat, table, arrays = fold_mobilenetv2(sd)
        arch, dims = ARCH_MOBILENETV2, [stem_c, len(table), 0, 0, feat]
    elif "conv1.conv.weight" in sd and "conv5.conv.weight" in sd and "fc.0.weight" in sd:
        dims, arrays = fold_osnet(sd)
        arch = ARCH_OSNET
    else:
        raise ValueError("only OSNet and MobileNetV2 state dicts are implemented on the B200 ReID path")
    # every tensor starts on a 16-byte boundary (the kernels read weights as float4)
    padded = []
    for a in arrays:
        flat = np.asarray(a, dtype=np.float32).ravel()
        padded.append(np.pad(flat, (0, (-flat.size) % 4)))
    payload = np.concatenate(padded)
    header = [MAGIC, VERSION, arch, *dims, int(payload.size)] + [0] * (16 - 9)
    out_path = Path(out_path)
    tmp = out_path.with_suffix(out_path.suffix + ".tmp")
    with open(tmp, "wb") as f:
        f.write(struct.pack("<16i", *header))
        for row in table:
            f.write(struct.pack("<4i", *row))
        f.write(payload.tobytes())
    tmp.replace(out_path)
    return out_path


def read_blob(path):
    raw = Path(path).read_bytes()
    header = struct.unpack("<16i", raw[:64])
    if header[0] != MAGIC or header[1] != VERSION:
        raise ValueError("not a .b200reid blob")
    off = 64
    if header[2] == ARCH_MOBILENETV2:
        off += 16 * header[4]
    payload = np.frombuffer(raw[off:], dtype=np.float32)
    assert payload.size == header[8]
    return header, payload


def read_block_table(path):
    raw = Path(path).read_bytes()
    header = struct.unpack("<16i", raw[:64])
    if header[2] != ARCH_MOBILENETV2:
        return []
    return [struct.unpack("<4i", raw[64 + 16 * i: 80 + 16 * i]) for i in range(header[4])]


def blob_digest(path) -> str:
    return hashlib.sha256(Path(path).read_bytes()).hexdigest()[:16]

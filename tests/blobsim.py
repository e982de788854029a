"""Walk a `.b200reid` blob exactly as csrc/reid_model.cu does and evaluate the folded network with torch ops
(NHWC, float32).  Test infrastructure: validates boxmot_b200/weights.py (BN folding, layouts, ordering)
against the oracle without a GPU, and documents the arithmetic each CUDA kernel implements."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from boxmot_b200.weights import read_blob

LIGHTS = (1, 2, 3, 4)  # branch depths


class _Cursor:
    def __init__(self, payload):
        self.p = torch.from_numpy(payload.copy())
        self.o = 0

    def take(self, *shape):
        n = int(np.prod(shape))
        t = self.p[self.o:self.o + n].view(*shape)
        self.o += (n + 3) // 4 * 4  # tensors are padded to 16 bytes
        return t


def _pw(x, w, b, relu):
    y = x @ w + b
    return F.relu(y) if relu else y


def _dw3(x, w9c, b):
    # x (N,H,W,C); depthwise 3x3, zero padding 1, weights [9][C] (tap = kh*3+kw)
    n, h, wd, c = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    acc = torch.zeros_like(x)
    for kh in range(3):
        for kw in range(3):
            acc = acc + xp[:, kh:kh + h, kw:kw + wd, :] * w9c[kh * 3 + kw]
    return F.relu(acc + b)


def _dw3s(x, w9c, b, stride):
    n, h, wd, c = x.shape
    y = F.conv2d(x.permute(0, 3, 1, 2), w9c.t().reshape(c, 1, 3, 3).contiguous(), b, stride=stride, padding=1, groups=c)
    return F.relu6(y).permute(0, 2, 3, 1).contiguous()


@torch.no_grad()
def blob_forward_mobilenetv2(blob_path, x_nhwc):
    from boxmot_b200.weights import read_block_table

    header, payload = read_blob(blob_path)
    table = read_block_table(blob_path)
    stem_c, feat = header[3], header[7]
    p4 = lambda n: (n + 3) // 4 * 4
    cur = _Cursor(payload)
    w, b = cur.take(27, p4(stem_c)), cur.take(p4(stem_c))
    wt = w.view(3, 3, 3, p4(stem_c)).permute(3, 2, 0, 1).contiguous()
    x = F.relu6(F.conv2d(x_nhwc.permute(0, 3, 1, 2), wt, b, stride=2, padding=1)).permute(0, 2, 3, 1).contiguous()
    for cin, cout, t, stride in table:
        mid = cin * t
        e = F.relu6(x @ cur.take(p4(cin), p4(mid)) + cur.take(p4(mid)))
        d = _dw3s(e, cur.take(9, p4(mid)), cur.take(p4(mid)), stride)
        y = d @ cur.take(p4(mid), p4(cout)) + cur.take(p4(cout))
        x = x + y if (stride == 1 and cin == cout) else y
    last = table[-1][1]
    x = F.relu6(x @ cur.take(p4(last), p4(feat)) + cur.take(p4(feat)))
    assert cur.o == payload.size
    return x.mean(dim=(1, 2))[:, :feat]


@torch.no_grad()
def blob_forward(blob_path, x_nhwc: torch.Tensor, return_stages=False):
    """x_nhwc (N,256,128,3) float32 normalised RGB -> (N, feat) un-normalised embedding."""
    header, payload = read_blob(blob_path)
    if header[2] == 2:
        return blob_forward_mobilenetv2(blob_path, x_nhwc)
    c = list(header[3:7])
    feat = header[7]
    cur = _Cursor(payload)
    stages = {}
    w, b = cur.take(147, c[0]), cur.take(c[0])
    xn = x_nhwc.permute(0, 3, 1, 2)
    wt = w.view(7, 7, 3, c[0]).permute(3, 2, 0, 1).contiguous()
    x = F.relu(F.conv2d(xn, wt, b, stride=2, padding=3))
    x = F.max_pool2d(x, 3, stride=2, padding=1).permute(0, 2, 3, 1).contiguous()
    stages["pool"] = x
    for s in range(3):
        for j in range(2):
            cin = c[s] if j == 0 else c[s + 1]
            cout = c[s + 1]
            mid, hid = cout // 4, cout // 64
            x1 = _pw(x, cur.take(cin, mid), cur.take(mid), True)
            branches = []
            for depth in LIGHTS:
                y = x1
                for _ in range(depth):
                    wpw, wdw, bb = cur.take(mid, mid), cur.take(9, mid), cur.take(mid)
                    y = _dw3(y @ wpw, wdw, bb)
                branches.append(y)
            w1, b1, w2, b2 = cur.take(mid, hid), cur.take(hid), cur.take(hid, mid), cur.take(mid)
            x2 = 0
            for y in branches:
                g = torch.sigmoid(F.relu(y.mean(dim=(1, 2)) @ w1 + b1) @ w2 + b2)
                x2 = x2 + y * g[:, None, None, :]
            if cin != cout:
                wc, bc = cur.take(mid + cin, cout), cur.take(cout)
                x = F.relu(torch.cat([x2, x], dim=-1) @ wc + bc)
            else:
                wc, bc = cur.take(mid, cout), cur.take(cout)
                x = F.relu(x2 @ wc + bc + x)
            stages[f"conv{s + 2}.{j}"] = x
        if s < 2:
            x = _pw(x, cur.take(c[s + 1], c[s + 1]), cur.take(c[s + 1]), True)
            n, h, wd, ch = x.shape
            x = x.view(n, h // 2, 2, wd // 2, 2, ch).mean(dim=(2, 4))
            stages[f"conv{s + 2}.2"] = x
    x = _pw(x, cur.take(c[3], c[3]), cur.take(c[3]), True)
    v = x.mean(dim=(1, 2))
    v = F.relu(v @ cur.take(c[3], feat) + cur.take(feat))
    assert cur.o == payload.size
    return (v, stages) if return_stages else v

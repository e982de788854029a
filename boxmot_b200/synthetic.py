"""Seeded synthetic inputs for benchmarks and tests: the reference's own throughput-sweep detection
generator and random-init OSNet weights.  Data generators only -- no arithmetic of the tracked path.

`bench_stream` restates /root/reference/tests/performance/benchmark_fps.py:60-94 (`_make_random_dets` +
`_jitter_dets`, seed 42 + n_dets (+ 1000 * stream), fixed random uint8 frame) -- SURVEY.md section 8(d).
`make_osnet_state` builds a state dict with the reference's parameter names (reid/backbones/osnet.py) because the
container has no pretrained checkpoints; BatchNorm statistics are randomised so folding is
exercised, and gains are chosen so activations stay O(1-10) through the network.
"""
from __future__ import annotations

import numpy as np

OSNET_ARCHS = {
    "osnet_x0_25": (16, 64, 96, 128),
    "osnet_x0_5": (32, 128, 192, 256),
    "osnet_x0_75": (48, 192, 288, 384),
    "osnet_x1_0": (64, 256, 384, 512),
}
BRANCH_DEPTHS = (("conv2a", 1), ("conv2b", 2), ("conv2c", 3), ("conv2d", 4))


def bench_image(hw=(720, 1280), seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 255, size=(hw[0], hw[1], 3), dtype=np.uint8)


def bench_stream(n_dets: int, n_frames: int, hw=(720, 1280), stream: int = 0):
    """Returns (image, [dets_f32 (n,6)] * n_frames) exactly as the reference benchmark would feed them."""
    h, w = hw
    rng = np.random.default_rng(42 + n_dets + 1000 * stream)
    img = rng.integers(0, 255, size=(h, w, 3), dtype=np.uint8)
    cx = rng.uniform(80, w - 80, size=n_dets)
    cy = rng.uniform(80, h - 80, size=n_dets)
    bw = rng.uniform(40, 100, size=n_dets)
    bh = rng.uniform(80, 200, size=n_dets)
    x1 = np.clip(cx - bw / 2, 0, w - 1)
    y1 = np.clip(cy - bh / 2, 0, h - 1)
    x2 = np.clip(cx + bw / 2, 1, w)
    y2 = np.clip(cy + bh / 2, 1, h)
    conf = rng.uniform(0.55, 0.95, size=n_dets)
    cls = np.zeros(n_dets, dtype=np.float32)
    base = np.stack([x1, y1, x2, y2, conf, cls], axis=1).astype(np.float32)
    frames = []
    for _ in range(n_frames):
        out = base.copy()
        dx = rng.normal(0.0, 4.0, size=n_dets).astype(np.float32)
        dy = rng.normal(0.0, 4.0, size=n_dets).astype(np.float32)
        out[:, 0] = np.clip(out[:, 0] + dx, 0, w - 1)
        out[:, 2] = np.clip(out[:, 2] + dx, 1, w)
        out[:, 1] = np.clip(out[:, 1] + dy, 0, h - 1)
        out[:, 3] = np.clip(out[:, 3] + dy, 1, h)
        frames.append(out)
    return img, frames



def make_osnet_state(arch: str = "osnet_x0_25", seed: int = 0, feature_dim: int = 512, num_classes: int = 1041):
    import torch

    ch = OSNET_ARCHS[arch]
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k, groups=1, gain=1.0):
        fan_in = (ci // groups) * k * k
        sd[name + ".weight"] = torch.randn(co, ci // groups, k, k, generator=g) * (gain / fan_in) ** 0.5

    def bn(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    def light(name, c):
        conv(name + ".conv1", c, c, 1)
        conv(name + ".conv2", c, c, 3, groups=c)
        bn(name + ".bn", c)

    def osblock(name, cin, cout):
        mid = cout // 4
        conv(name + ".conv1.conv", mid, cin, 1)
        bn(name + ".conv1.bn", mid)
        light(name + ".conv2a", mid)
        for br, depth in BRANCH_DEPTHS[1:]:
            for k in range(depth):
                light(f"{name}.{br}.{k}", mid)
        hid = mid // 16
        conv(name + ".gate.fc1", hid, mid, 1)
        sd[name + ".gate.fc1.bias"] = 0.1 * torch.randn(hid, generator=g)
        conv(name + ".gate.fc2", mid, hid, 1)
        sd[name + ".gate.fc2.bias"] = 0.1 * torch.randn(mid, generator=g)
        conv(name + ".conv3.conv", cout, mid, 1, gain=0.1)
        bn(name + ".conv3.bn", cout)
        if cin != cout:
            conv(name + ".downsample.conv", cout, cin, 1)
            bn(name + ".downsample.bn", cout)

    conv("conv1.conv", ch[0], 3, 7)
    bn("conv1.bn", ch[0])
    for s, (cin, cout) in enumerate(((ch[0], ch[1]), (ch[1], ch[2]), (ch[2], ch[3]))):
        stage = f"conv{s + 2}"
        osblock(f"{stage}.0", cin, cout)
        osblock(f"{stage}.1", cout, cout)
        if s < 2:
            conv(f"{stage}.2.0.conv", cout, cout, 1)
            bn(f"{stage}.2.0.bn", cout)
    conv("conv5.conv", ch[3], ch[3], 1)
    bn("conv5.bn", ch[3])
    sd["fc.0.weight"] = 0.05 * torch.randn(feature_dim, ch[3], generator=g)
    sd["fc.0.bias"] = 0.05 * torch.randn(feature_dim, generator=g)
    bn("fc.1", feature_dim)
    sd["classifier.weight"] = 0.01 * torch.randn(num_classes, feature_dim, generator=g)
    sd["classifier.bias"] = torch.zeros(num_classes)
    return sd


MOBILENETV2_LAYERS = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                      (6, 320, 1, 1))  # (expansion t, base channels c, repeats n, first stride s), mobilenetv2.py:91-99


def mobilenetv2_blocks(width_mult: float = 1.4):
    """[(cin, cout, t, stride)] for every Bottleneck, plus stem channels and feature dim (mobilenetv2.py:86-102)."""
    cin = int(32 * width_mult)
    stem = cin
    blocks = []
    for t, c, n, s in MOBILENETV2_LAYERS:
        cout = int(c * width_mult)
        for i in range(n):
            blocks.append((cin, cout, t, s if i == 0 else 1))
            cin = cout
    feat = int(1280 * width_mult) if width_mult > 1 else 1280
    return stem, blocks, feat


def make_mobilenetv2_state(width_mult: float = 1.4, seed: int = 0, num_classes: int = 1041):
    """Seeded state dict with the reference's MobileNetV2 parameter names (reid/backbones/mobilenetv2.py)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k, groups=1, gain=1.0):
        fan_in = (ci // groups) * k * k
        sd[name + ".weight"] = torch.randn(co, ci // groups, k, k, generator=g) * (gain / fan_in) ** 0.5

    def bn(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    stem, blocks, feat = mobilenetv2_blocks(width_mult)
    conv("conv1.conv", stem, 3, 3, gain=2.0)
    bn("conv1.bn", stem)
    stage, idx, prev = 2, 0, None
    layer_of = []
    for t, c, n, s in MOBILENETV2_LAYERS:
        for i in range(n):
            layer_of.append((stage, i))
        stage += 1
    for (cin, cout, t, stride), (st, i) in zip(blocks, layer_of):
        name = f"conv{st}.{i}"
        mid = cin * t
        conv(name + ".conv1.conv", mid, cin, 1, gain=2.0)
        bn(name + ".conv1.bn", mid)
        conv(name + ".dwconv2.conv", mid, mid, 3, groups=mid, gain=2.0)
        bn(name + ".dwconv2.bn", mid)
        conv(name + ".conv3.0", cout, mid, 1, gain=0.5)
        bn(name + ".conv3.1", cout)
    conv("conv9.conv", feat, blocks[-1][1], 1, gain=2.0)
    bn("conv9.bn", feat)
    sd["classifier.weight"] = 0.01 * torch.randn(num_classes, feat, generator=g)
    sd["classifier.bias"] = torch.zeros(num_classes)
    return sd


def cohort_stream(n_objects=2048, cohorts=4, frames=10, hw=(1080, 1920), seed=3, dim=512, conf_lo=0.4):
    """BASELINE config 3 generator: `n_objects` fixed objects in `cohorts` cohorts, cohort f mod cohorts is
    visible on frame f (every track is re-observed every `cohorts` frames < max_age) -> ~n_objects live tracks and
    n_objects / cohorts detections per frame.  Returns (dets per frame, unit appearance rows per frame)."""
    rng = np.random.default_rng(seed)
    h, w = hw
    cx, cy = rng.uniform(40, w - 40, n_objects), rng.uniform(60, h - 60, n_objects)
    bw, bh = rng.uniform(20, 50, n_objects), rng.uniform(40, 100, n_objects)
    conf = rng.uniform(conf_lo, 0.95, n_objects)   # conf_lo above the tracker's det_thresh: every object is tracked
    proto = np.abs(rng.normal(size=(n_objects, dim))).astype(np.float32)
    dets, embs = [], []
    for f in range(frames):
        idx = np.arange(f % cohorts, n_objects, cohorts)
        jx, jy = rng.normal(0, 2.0, idx.size), rng.normal(0, 2.0, idx.size)
        d = np.stack([cx[idx] + jx - bw[idx] / 2, cy[idx] + jy - bh[idx] / 2, cx[idx] + jx + bw[idx] / 2,
                      cy[idx] + jy + bh[idx] / 2, conf[idx], np.zeros(idx.size)], 1).astype(np.float32)
        e = np.maximum(proto[idx] + 0.3 * rng.normal(size=(idx.size, dim)).astype(np.float32), 0)
        dets.append(d)
        embs.append((e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32))
    return dets, embs


def camera_pan_sequence(n_frames: int = 16, hw=(360, 640), n_objects: int = 24, seed: int = 5, max_step: float = 4.0,
                        dim: int = 0):
    """Moving-camera test input: a smooth textured canvas seen through a window that pans by a few whole
    pixels per frame (plus per-frame sensor noise), and detections of objects that stand still on the canvas -- so they
    move in the image by exactly the camera motion.  Pure numpy, seeded.  Returns (frames [H,W,3] uint8 BGR, dets (n,6)
    float32 per frame, window offsets (n_frames, 2) int, embeddings per frame or None)."""
    h, w = hw
    rng = np.random.default_rng(seed)
    margin = int(max_step * n_frames) + 8
    H, W = h + 2 * margin, w + 2 * margin
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    canvas = np.zeros((H, W, 3), np.float32)
    def smooth_field(cell):   # random grid of `cell`-pixel cells, bilinearly interpolated: smooth at the 0.15 scale
        gh, gw = H // cell + 3, W // cell + 3
        g = rng.random((gh, gw)).astype(np.float32)
        fy, fx = yy / cell, xx / cell
        y0, x0 = fy.astype(np.int64), fx.astype(np.int64)
        ay, ax = fy - y0, fx - x0
        return (g[y0, x0] * (1 - ay) * (1 - ax) + g[y0, x0 + 1] * (1 - ay) * ax
                + g[y0 + 1, x0] * ay * (1 - ax) + g[y0 + 1, x0 + 1] * ay * ax)

    for c in range(3):
        canvas[..., c] = smooth_field(56) + 0.5 * smooth_field(28) + 0.08 * rng.random((H, W)).astype(np.float32)
    canvas -= canvas.min()
    canvas = canvas / canvas.max() * 215.0 + 20.0
    ox, oy = float(margin), float(margin)
    offs, frames, dets, embs = [], [], [], []
    cx = rng.uniform(margin + 40, margin + w - 40, n_objects)
    cy = rng.uniform(margin + 40, margin + h - 40, n_objects)
    bw = rng.uniform(18, 46, n_objects)
    bh = rng.uniform(36, 90, n_objects)
    protos = np.abs(rng.normal(size=(n_objects, max(dim, 1)))).astype(np.float32)
    vx, vy = rng.uniform(-max_step, max_step, 2)
    for f in range(n_frames):
        if f % 5 == 0:
            vx, vy = rng.uniform(-max_step, max_step, 2)
        ox = float(np.clip(ox + vx, 2, 2 * margin - 2))
        oy = float(np.clip(oy + vy, 2, 2 * margin - 2))
        ix, iy = int(round(ox)), int(round(oy))
        offs.append((ix, iy))
        win = canvas[iy:iy + h, ix:ix + w] + rng.normal(0.0, 1.5, (h, w, 3)).astype(np.float32)
        frames.append(np.clip(np.rint(win), 0, 255).astype(np.uint8))
        keep = rng.random(n_objects) > 0.1
        x1 = cx - bw / 2 - ix + rng.normal(0, 0.4, n_objects)
        y1 = cy - bh / 2 - iy + rng.normal(0, 0.4, n_objects)
        d = np.stack([x1, y1, x1 + bw, y1 + bh, rng.uniform(0.55, 0.95, n_objects), np.zeros(n_objects)], 1)
        vis = keep & (d[:, 0] > 0) & (d[:, 1] > 0) & (d[:, 2] < w) & (d[:, 3] < h)
        dets.append(d[vis].astype(np.float32))
        if dim:
            e = np.maximum(protos[vis] + 0.3 * rng.normal(size=(int(vis.sum()), dim)).astype(np.float32), 0.0)
            embs.append((e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32))
    return frames, dets, np.asarray(offs), (embs if dim else None)

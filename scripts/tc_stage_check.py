"""Stage-by-stage comparison of the tensor-core OSNet path with the oracle (run on the GPU box):
prints max |device - oracle| / max |oracle| for every tap the plan exposes -- pool (2), conv1 outputs (100 + block),
branch outputs (200 + block), block / transition outputs (3..10), conv5 (11) -- and the embedding error.
TEST-SIDE TOOL: imports the oracle; nothing here is product code."""
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import reid as orid  # noqa: E402
from boxmot_b200.reid import B200ReID  # noqa: E402
from boxmot_b200.weights import export_blob  # noqa: E402
from boxmot_b200.synthetic import BRANCH_DEPTHS  # noqa: E402


def block_parts(sd, name, x):
    x1 = F.relu(orid._bn(sd, name + ".conv1.bn", F.conv2d(x, sd[name + ".conv1.conv.weight"])))
    branches = [orid._light(sd, name + ".conv2a", x1)]
    for br, depth in BRANCH_DEPTHS[1:]:
        y = x1
        for k in range(depth):
            y = orid._light(sd, f"{name}.{br}.{k}", y)
        branches.append(y)
    return x1, branches


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "osnet_x0_25"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    sd = orid.make_osnet_state(arch, seed=11)
    tmp = Path(tempfile.mkdtemp())
    reid = B200ReID(export_blob(sd, tmp / "m.b200reid"))
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, size=(360, 640, 3), dtype=np.uint8)
    boxes = np.array([[10, 20, 90, 200], [300, 100, 380, 330], [-20, -10, 60, 100], [600, 300, 700, 400],
                      [100.5, 50.5, 101.4, 52.2]], np.float32)
    if n > 5:
        cx, cy = rng.uniform(0, 640, n - 5), rng.uniform(0, 360, n - 5)
        w, h = rng.uniform(20, 120, n - 5), rng.uniform(40, 240, n - 5)
        boxes = np.concatenate([boxes, np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)])
    with torch.no_grad():
        x0 = orid.get_crops(boxes, img)
        emb, stages = orid.osnet_forward(sd, x0, return_stages=True)
        want = {1: nhwc(stages["stem"]), 2: nhwc(stages["pool"])}
        names = ["conv2.0", "conv2.1", "conv2.2", "conv3.0", "conv3.1", "conv3.2", "conv4.0", "conv4.1"]
        for i, nm in enumerate(names):
            want[3 + i] = nhwc(stages[nm])
        x = stages["pool"]
        bi = 0
        for s in range(3):
            for j in range(2):
                name = f"conv{s + 2}.{j}"
                x1, br = block_parts(sd, name, x)
                mid = x1.shape[1]
                midp = (mid + 15) // 16 * 16
                pad = lambda t: F.pad(t, (0, 0, 0, 0, 0, midp - mid))
                want[100 + bi] = nhwc(pad(x1))
                want[200 + bi] = nhwc(torch.cat([pad(b) for b in br], 1))
                x = stages[name]
                bi += 1
            if s < 2:
                x = stages[f"conv{s + 2}.2"]
        c5 = F.relu(orid._bn(sd, "conv5.bn", F.conv2d(x, sd["conv5.conv.weight"])))
        want[11] = nhwc(c5)
    g50 = reid.debug_stage(boxes, img, 50).reshape(len(boxes), 256, 128, 3)
    w50 = orid.crop_boxes(boxes, img).astype(np.float32)
    print("stage   50 (fused resize, uint8 RGB): mismatching values", int((g50 != w50).sum()), "of", g50.size)
    order = [1, 2, 100, 200, 3, 101, 201, 4, 5, 102, 202, 6, 103, 203, 7, 8, 104, 204, 9, 105, 205, 10, 11]
    worst = 0.0
    for st in order:
        try:
            g = reid.debug_stage(boxes, img, st)
        except Exception as e:  # taps of fused launches do not exist
            print(f"stage {st:4d}: n/a ({str(e)[:60]})")
            continue
        w = want[st].reshape(len(boxes), -1)
        if g.shape != w.shape:
            print(f"stage {st:4d}: shape {g.shape} vs oracle {w.shape}")
            continue
        err = np.abs(g - w).max() / max(1e-30, np.abs(w).max())
        bad = np.argwhere(np.abs(g - w) > 1e-3 * np.abs(w).max())
        extra = ""
        if len(bad):
            c = want[st].shape[-1]
            hw = want[st].shape[1] * want[st].shape[2]
            wd = want[st].shape[2]
            b0 = bad[0]
            p, ch = divmod(int(b0[1]), c)
            extra = (f"  first bad: crop {b0[0]} y {p // wd} x {p % wd} ch {ch} got {g[b0[0], b0[1]]:.5f} want {w[b0[0], b0[1]]:.5f};"
                     f" bad {len(bad)}/{g.size}, crops {sorted(set(bad[:, 0].tolist()))[:6]},"
                     f" chans {sorted(set((bad[:, 1] % c).tolist()))[:12]}")
        print(f"stage {st:4d}: rel-to-max err {err:.3e}{extra}")
        if st >= 3 and st < 100:
            worst = max(worst, err)
    feats = reid.get_features(boxes, img)
    e = emb.numpy()
    e = e / np.linalg.norm(e, axis=1, keepdims=True)
    rel = (np.abs(feats - e).max(1) / np.abs(e).max(1)).max()
    print(f"embedding: max_i |d| / |e|inf = {rel:.3e} (bound 1e-4); worst block stage {worst:.3e}")


if __name__ == "__main__":
    main()

"""Oracle restatement of the ReID path: crop staging + OSNet forward -- TEST INFRASTRUCTURE ONLY.

Follows (relative to /root/reference/boxmot):
  * reid/backends/base_backend.py:148-195  get_crops: round -> clip -> slice (blank 256x128 when empty) ->
    cv2.resize INTER_LINEAR to (128 w, 256 h) -> BGR2RGB -> /255 -> (x - mean) / std, float32 NCHW
  * reid/core/preprocessing.py:12-18       resize
  * reid/backends/base_backend.py:197-207  get_features: forward, then row-wise L2 normalisation
  * reid/backbones/osnet.py:27-260,380-405 ConvLayer / Conv1x1 / Conv1x1Linear / LightConv3x3 / ChannelGate
    (one gate module shared by the four branches) / OSBlock / OSNet.forward in eval mode
  * reid/backbones/osnet.py:488,533        osnet_x1_0 (64/256/384/512) and osnet_x0_25 (16/64/96/128)
`resize_linear_u8` restates OpenCV's 8-bit INTER_LINEAR fixed-point path (third-party, opencv 4.13 in this
image): 11-bit coefficients from float32 phase, horizontal pass in int32, vertical
(((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; x phases are clamped at the borders, y rows are clipped
at fetch.  It is pinned bit-for-bit against cv2.resize in tests/test_oracle_reid.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
INPUT_HW = (256, 128)

from boxmot_b200.synthetic import BRANCH_DEPTHS, OSNET_ARCHS  # noqa: E402


# ----------------------------------------------------------------------------------------------------
# crop staging
# ----------------------------------------------------------------------------------------------------
def _linear_coeffs(dst_n: int, src_n: int, clamp: bool):
    inv_scale = float(dst_n) / float(src_n)
    scale = 1.0 / inv_scale
    idx = np.zeros(dst_n, np.int64)
    a0 = np.zeros(dst_n, np.int64)
    a1 = np.zeros(dst_n, np.int64)
    for d in range(dst_n):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= src_n - 1:
                s, f = src_n - 1, np.float32(0)
        idx[d] = s
        a0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
    return idx, a0, a1


def resize_linear_u8(src: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    if (sh, sw) == (dst_h, dst_w):
        return src.copy()
    xi, xa0, xa1 = _linear_coeffs(dst_w, sw, True)
    yi, ya0, ya1 = _linear_coeffs(dst_h, sh, False)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    hor = s[:, xi, :] * xa0[None, :, None] + s[:, x1, :] * xa1[None, :, None]
    y0 = np.clip(yi, 0, sh - 1)
    y1 = np.clip(yi + 1, 0, sh - 1)
    b0 = ya0[:, None, None]
    b1 = ya1[:, None, None]
    out = (((b0 * (hor[y0] >> 4)) >> 16) + ((b1 * (hor[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


IMAGENET_MEAN_BGR = (104, 116, 124)


def resize_pad_u8(src: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """reid/core/preprocessing.py:21-45: aspect-preserving resize (scale = min(W / w, H / h), new = int(size * scale)),
    centred, constant ImageNet-mean border (BGR order: the channel flip comes later)."""
    h, w = src.shape[:2]
    scale = min(dst_w / w, dst_h / h)
    new_w, new_h = int(w * scale), int(h * scale)
    resized = resize_linear_u8(src, new_h, new_w)
    top, left = (dst_h - new_h) // 2, (dst_w - new_w) // 2
    out = np.empty((dst_h, dst_w, 3), np.uint8)
    out[:] = np.array(IMAGENET_MEAN_BGR, np.uint8)
    out[top:top + new_h, left:left + new_w] = resized
    return out


def crop_boxes(xyxys: np.ndarray, img: np.ndarray, preprocess: str = "resize"):
    """uint8 RGB crops (N,256,128,3) exactly as get_crops stages them before the float conversion."""
    h, w = img.shape[:2]
    xyxys = np.asarray(xyxys, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((len(xyxys), INPUT_HW[0], INPUT_HW[1], 3), np.uint8)
    for i, box in enumerate(xyxys):
        x1, y1, x2, y2 = box.round().astype("int")
        cx1, cy1 = max(0, x1), max(0, y1)
        cx2, cy2 = min(w, x2), min(h, y2)
        if cx2 > cx1 and cy2 > cy1:
            fn = resize_pad_u8 if preprocess == "resize_pad" else resize_linear_u8
            crop = fn(img[cy1:cy2, cx1:cx2], INPUT_HW[0], INPUT_HW[1])
        else:
            crop = np.zeros((INPUT_HW[0], INPUT_HW[1], 3), np.uint8)
        out[i] = crop[:, :, ::-1]
    return out


def get_crops(xyxys: np.ndarray, img: np.ndarray, preprocess: str = "resize") -> torch.Tensor:
    """float32 NCHW network input (N,3,256,128)."""
    u8 = crop_boxes(xyxys, img, preprocess)
    x = torch.from_numpy(u8).to(torch.float32).permute(0, 3, 1, 2).contiguous()
    x = x / 255.0
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return (x - mean) / std


from boxmot_b200.synthetic import make_osnet_state  # noqa: E402,F401  (seeded weights; no arithmetic of the path)


def detect_osnet_arch(sd) -> str:
    c0 = sd["conv1.conv.weight"].shape[0]
    c3 = sd["conv5.conv.weight"].shape[0]
    for name, ch in OSNET_ARCHS.items():
        if ch[0] == c0 and ch[3] == c3:
            return name
    raise ValueError("state dict is not an OSNet of a known width")


# ----------------------------------------------------------------------------------------------------
# functional forward (eval mode), float32
# ----------------------------------------------------------------------------------------------------
def _bn(sd, name, x, eps=1e-5):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=eps)


def _light(sd, name, x):
    c = x.shape[1]
    x = F.conv2d(x, sd[name + ".conv1.weight"])
    x = F.conv2d(x, sd[name + ".conv2.weight"], padding=1, groups=c)
    return F.relu(_bn(sd, name + ".bn", x))


def _gate(sd, name, x):
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.relu(F.conv2d(g, sd[name + ".fc1.weight"], sd[name + ".fc1.bias"]))
    g = torch.sigmoid(F.conv2d(g, sd[name + ".fc2.weight"], sd[name + ".fc2.bias"]))
    return x * g


def _osblock(sd, name, x):
    identity = x
    x1 = F.relu(_bn(sd, name + ".conv1.bn", F.conv2d(x, sd[name + ".conv1.conv.weight"])))
    branches = [_light(sd, name + ".conv2a", x1)]
    for br, depth in BRANCH_DEPTHS[1:]:
        y = x1
        for k in range(depth):
            y = _light(sd, f"{name}.{br}.{k}", y)
        branches.append(y)
    x2 = (_gate(sd, name + ".gate", branches[0]) + _gate(sd, name + ".gate", branches[1])
          + _gate(sd, name + ".gate", branches[2]) + _gate(sd, name + ".gate", branches[3]))
    x3 = _bn(sd, name + ".conv3.bn", F.conv2d(x2, sd[name + ".conv3.conv.weight"]))
    if (name + ".downsample.conv.weight") in sd:
        identity = _bn(sd, name + ".downsample.bn", F.conv2d(identity, sd[name + ".downsample.conv.weight"]))
    return F.relu(x3 + identity)


@torch.no_grad()
def osnet_forward(sd, x: torch.Tensor, return_stages: bool = False):
    """x (N,3,256,128) float32 -> (N, feature_dim) un-normalised embedding (fc output, eval mode)."""
    stages = {}
    x = F.relu(_bn(sd, "conv1.bn", F.conv2d(x, sd["conv1.conv.weight"], stride=2, padding=3)))
    stages["stem"] = x
    x = F.max_pool2d(x, 3, stride=2, padding=1)
    stages["pool"] = x
    for s in range(3):
        stage = f"conv{s + 2}"
        x = _osblock(sd, f"{stage}.0", x)
        stages[f"{stage}.0"] = x
        x = _osblock(sd, f"{stage}.1", x)
        stages[f"{stage}.1"] = x
        if s < 2:
            x = F.relu(_bn(sd, f"{stage}.2.0.bn", F.conv2d(x, sd[f"{stage}.2.0.conv.weight"])))
            x = F.avg_pool2d(x, 2, stride=2)
            stages[f"{stage}.2"] = x
    x = F.relu(_bn(sd, "conv5.bn", F.conv2d(x, sd["conv5.conv.weight"])))
    v = F.adaptive_avg_pool2d(x, 1).flatten(1)
    v = F.linear(v, sd["fc.0.weight"], sd["fc.0.bias"])
    v = F.relu(_bn(sd, "fc.1", v))
    return (v, stages) if return_stages else v


def is_mobilenetv2(sd) -> bool:
    return "conv9.conv.weight" in sd


@torch.no_grad()
def mobilenetv2_forward(sd, x: torch.Tensor, return_stages: bool = False):
    """reid/backbones/mobilenetv2.py:19-41 (ConvBlock = conv + BN + ReLU6), :43-77 (Bottleneck), :168-198 (forward):
    x (N,3,256,128) -> (N, feature_dim) global-average-pooled conv9 map (eval mode, no fc)."""
    stages = {}

    def block(name, y, k, s=1, p=0, g=1):
        return F.relu6(_bn(sd, name + ".bn", F.conv2d(y, sd[name + ".conv.weight"], stride=s, padding=p, groups=g)))

    x = block("conv1", x, 3, s=2, p=1)
    stages["conv1"] = x
    stage = 2
    while f"conv{stage}.0.conv1.conv.weight" in sd:
        i = 0
        while f"conv{stage}.{i}.conv1.conv.weight" in sd:
            name = f"conv{stage}.{i}"
            mid = sd[name + ".conv1.conv.weight"].shape[0]
            cin = sd[name + ".conv1.conv.weight"].shape[1]
            cout = sd[name + ".conv3.0.weight"].shape[0]
            # the first Bottleneck of a stage carries the stage stride (mobilenetv2.py:104-110)
            from boxmot_b200.synthetic import MOBILENETV2_LAYERS
            stride = MOBILENETV2_LAYERS[stage - 2][3] if i == 0 else 1
            m = block(name + ".conv1", x, 1)
            m = block(name + ".dwconv2", m, 3, s=stride, p=1, g=mid)
            m = _bn(sd, name + ".conv3.1", F.conv2d(m, sd[name + ".conv3.0.weight"]))
            x = x + m if (stride == 1 and cin == cout) else m
            stages[name] = x
            i += 1
        stage += 1
    x = block("conv9", x, 1)
    v = F.adaptive_avg_pool2d(x, 1).flatten(1)
    return (v, stages) if return_stages else v


def backbone_forward(sd, x):
    return mobilenetv2_forward(sd, x) if is_mobilenetv2(sd) else osnet_forward(sd, x)


def get_features(sd, xyxys: np.ndarray, img: np.ndarray, preprocess: str = "resize") -> np.ndarray:
    """(N, D) float32 L2-normalised embeddings, as BaseModelBackend.get_features returns them."""
    xyxys = np.asarray(xyxys, dtype=np.float32)
    if xyxys.size == 0:
        return np.array([])
    feats = backbone_forward(sd, get_crops(xyxys, img, preprocess)).numpy()
    return feats / np.linalg.norm(feats, axis=-1, keepdims=True)


class OracleReID:
    """Minimal `reid_model` object for the oracle trackers (get_features only)."""

    def __init__(self, sd):
        self.sd = sd

    def get_features(self, xyxys, img):
        return get_features(self.sd, xyxys, img)

"""Seeded synthetic inputs for the hot path — shared by the tests, the
benchmark and the golden-fixture generator.  Everything is generated on the CPU from an
explicit seed so the same tensors can be fed to the CUDA operators and to the CPU oracle.

There are no datasets or checkpoints in this environment: features, heatmaps, keypoints and
the PoseResNet-101-shaped parameter list are all synthetic, of the reference's shapes.
"""
from __future__ import annotations

import numpy as np
import torch

# BASELINE.json configs: (batch per GPU, keypoints, sigma)
CONFIGS = {
    "C1": dict(name="RHD->H3D hand", batch=32, joints=21, sigma=2),
    "C2": dict(name="SURREAL->LSP human", batch=32, joints=16, sigma=2),
    "C3": dict(name="SURREAL->Human36M human", batch=32, joints=16, sigma=2),
    "C4": dict(name="SyntheticAnimal->TigDog animal", batch=64, joints=18, sigma=1.0),
    "C5": dict(name="kernel microbench", batch=256, joints=21, sigma=2),
}


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def vgg_features(n: int, seed: int, channels: int = 512, h: int = 32, w: int = 32):
    """(content, style) relu4_1-like features [n,C,h,w] fp32: post-ReLU statistics."""
    g = _gen(seed)
    content = torch.relu(torch.randn(n, channels, h, w, generator=g) * 1.0 + 0.2)
    style = torch.relu(torch.randn(n, channels, h, w, generator=g) * 2.0 + 0.5)
    return content, style


def heatmaps(b: int, k: int, seed: int, peak=(0.2, 1.1), h: int = 64, w: int = 64, noise: float = 0.05,
             sigma: float = 2.0) -> torch.Tensor:
    """[b,k,h,w] fp32 = noise*randn + one Gaussian bump with a uniform-random peak/position."""
    g = _gen(seed)
    hm = torch.randn(b, k, h, w, generator=g) * noise
    cx = torch.randint(0, w, (b, k, 1, 1), generator=g).float()
    cy = torch.randint(0, h, (b, k, 1, 1), generator=g).float()
    pk = torch.rand(b, k, 1, 1, generator=g) * (peak[1] - peak[0]) + peak[0]
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w)
    ys = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1)
    hm += pk * torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2.0 * sigma * sigma))
    return hm


def keypoints(b: int, k: int, seed: int, image: int = 256):
    """(joints float64 [b,k,2] in image pixels, vis float32 [b,k,1]): U[0,image)^2, 10 % invisible,
    5 % pushed out of bounds (negative or beyond the image)."""
    rng = np.random.RandomState(int(seed))
    joints = rng.uniform(0.0, image, size=(b, k, 2))
    vis = (rng.uniform(size=(b, k, 1)) >= 0.10).astype(np.float32)
    oob = rng.uniform(size=(b, k)) < 0.05
    shift = rng.choice([-1.5 * image, 1.5 * image], size=(b, k))
    axis = rng.randint(0, 2, size=(b, k))
    for a in (0, 1):
        sel = oob & (axis == a)
        joints[..., a][sel] += shift[sel]
    return joints, vis


def aug_params(b: int, seed: int, image: int = 256, degrees: float = 180.0, shear=(-30.0, 30.0),
               translate=(0.05, 0.05), scale=(0.6, 1.3), shear_y: bool = False):
    """The collated ``meta['aug_param_*']`` of one batch: ``[angle[b], [trans_x[b], trans_y[b]],
    [shear_x[b], shear_y[b]], scale[b]]`` (float64 / int64 tensors, as the DataLoader collates the
    Python floats / ints of lib/transforms/keypoint_detection.py:139), drawn like
    ``RandomAffineRotation.get_params`` (:397-412) with the trainers' default ranges
    (train_human.py:535-557) and inverted like ``affine()`` (:139)."""
    rng = np.random.RandomState(int(seed))
    angle = rng.uniform(-degrees, degrees, size=b)
    shx = rng.uniform(shear[0], shear[1], size=b)
    shy = rng.uniform(shear[0], shear[1], size=b) if shear_y else np.zeros(b)
    tx = np.round(rng.uniform(-translate[0] * image, translate[0] * image, size=b)).astype(np.int64)
    ty = np.round(rng.uniform(-translate[1] * image, translate[1] * image, size=b)).astype(np.int64)
    sc = rng.uniform(scale[0], scale[1], size=b)
    t = torch.from_numpy
    return [t(-angle), [t(-tx), t(-ty)], [t(-shx), t(-shy)], t(1.0 / sc)]


def adversarial_heatmaps(k: int = 4, h: int = 64, w: int = 64, seed: int = 7) -> torch.Tensor:
    """[10,k,h,w] fp32 planes that stress exact arg-max semantics: duplicated maxima (first index
    must win), all-zero, all-negative, NaN (NaN is the maximum, first NaN wins), +/-inf, -0.0/+0.0
    ties, maxima on rows/cols 0-1 (PCK validity boundary) and on the last cell."""
    g = _gen(seed)
    out = torch.randn(10, k, h, w, generator=g) * 0.1
    n = h * w
    flat = out.view(10, k, n)
    for j in range(k):
        # 0: duplicated maxima at two random cells
        a, b_ = sorted(torch.randint(0, n, (2,), generator=g).tolist())
        flat[0, j, a] = 3.0
        flat[0, j, b_] = 3.0
        # 1: all zero;  2: all negative
        flat[1, j] = 0.0
        flat[2, j] = -flat[2, j].abs() - 0.01
        # 3: NaNs (two of them) plus a large finite value
        pos = torch.randint(0, n, (3,), generator=g).tolist()
        flat[3, j, pos[0]] = float("nan")
        flat[3, j, pos[1]] = float("nan")
        flat[3, j, pos[2]] = 100.0
        # 4: +inf twice;  5: -inf everywhere except one finite cell
        flat[4, j, pos[0]] = float("inf")
        flat[4, j, pos[1]] = float("inf")
        flat[5, j] = float("-inf")
        flat[5, j, pos[2]] = -5.0
        # 6: -0.0 then +0.0 as the tied maximum among negatives
        flat[6, j] = -flat[6, j].abs() - 0.01
        flat[6, j, min(pos[0], pos[1])] = -0.0
        flat[6, j, max(pos[0], pos[1]) if pos[0] != pos[1] else (pos[0] + 1) % n] = 0.0
        # 7: maximum on the PCK validity boundary (x or y in {0, 1, 2})
        flat[7, j, (j % 3) * w + (2 - j % 3)] = 2.0
        # 8: maximum in the last cell;  9: constant positive plane (every cell ties)
        flat[8, j, n - 1] = 2.0
        flat[9, j] = 0.5
    return out


def pose_resnet_param_shapes(num_keypoints: int, layers=(3, 4, 23, 3)):
    """Parameter shapes of ``pose_resnet101(num_keypoints)`` in module order
    (lib/models/pose_resnet.py:59-112, lib/models/resnet.py): ResNet trunk without fc, three
    bias-free 4x4 deconvolutions + BN, and the 1x1 head.  323 tensors / 52 992 853 elements
    for K=21."""
    shapes = [(64, 3, 7, 7), (64,), (64,)]
    inplanes = 64
    for stage, blocks in enumerate(layers):
        planes = 64 * 2 ** stage
        for blk in range(blocks):
            shapes += [(planes, inplanes, 1, 1), (planes,), (planes,),
                       (planes, planes, 3, 3), (planes,), (planes,),
                       (planes * 4, planes, 1, 1), (planes * 4,), (planes * 4,)]
            if blk == 0:
                shapes += [(planes * 4, inplanes, 1, 1), (planes * 4,), (planes * 4,)]
            inplanes = planes * 4
    cin = inplanes
    for _ in range(3):
        shapes += [(cin, 256, 4, 4), (256,), (256,)]
        cin = 256
    shapes += [(num_keypoints, 256, 1, 1), (num_keypoints,)]
    return shapes


def parameter_list(shapes, seed: int, device="cpu", dtype=torch.float32, scale: float = 0.02):
    """One tensor per shape, ``randn * scale`` (each its own allocation, like real Parameters)."""
    g = _gen(seed)
    out = []
    for s in shapes:
        t = torch.randn(*s, generator=g) * scale
        out.append(t.to(device=device, dtype=dtype))
    return out

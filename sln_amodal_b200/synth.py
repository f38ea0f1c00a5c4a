"""Seeded synthetic inputs for the hot path.

Pure numpy; shared by tests/ and bench.py so both sides of every parity check and
every timing see the same tensors.  No datasets or weights exist offline, so these
generators define the workloads: COCOA/D2SA-shaped ROIs, RPN-like box clusters and
painter's-order instance label maps.
"""
from __future__ import annotations

import numpy as np

FPN_SIZES = {2: 256, 3: 128, 4: 64, 5: 32}   # P2..P5 side at a 1024^2 image (config.py:58)


def roi_boxes(n, seed=4321, image=1024.0, outside_frac=0.0, degenerate_frac=0.0, window=None):
    """Config-2 ROI generator: centres U(0,1), sqrt(area) log-uniform in [16,768] px of 1024,
    aspect log-uniform in [1/2,2], clipped to [0,1].  Returns f32[n,4] (y1,x1,y2,x2) normalised.

    outside_frac    : fraction of boxes shifted partially outside [0,1] (extrapolation set (i))
    degenerate_frac : fraction with y1==y2 (set (ii))
    window          : (cy, cx, side) normalised -- all boxes drawn inside that window (set (iii))
    """
    rng = np.random.default_rng(seed)
    s = np.exp(rng.uniform(np.log(16.0), np.log(768.0), n)) / image
    a = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    h = s * np.sqrt(a)
    w = s / np.sqrt(a)
    cy = rng.uniform(0, 1, n)
    cx = rng.uniform(0, 1, n)
    if window is not None:
        wy, wx, side = window
        h = np.minimum(h, side)
        w = np.minimum(w, side)
        cy = wy + (rng.uniform(-0.5, 0.5, n)) * (side - h)
        cx = wx + (rng.uniform(-0.5, 0.5, n)) * (side - w)
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1)
    b = np.clip(b, 0.0, 1.0)
    n_out = int(round(outside_frac * n))
    if n_out:
        ix = rng.choice(n, n_out, replace=False)
        b[ix] += rng.uniform(-0.3, 0.3, (n_out, 1)) * np.array([[1, 1, 1, 1]])
        b[ix, 2:] += rng.uniform(0.0, 0.2, (n_out, 2))
    n_deg = int(round(degenerate_frac * n))
    if n_deg:
        ix = rng.choice(n, n_deg, replace=False)
        b[ix, 2] = b[ix, 0]
    return b.astype(np.float32)


def fpn_level(boxes, image_hw=(1024, 1024)):
    """FPN level of each ROI, modals.py:53-64, in float64 numpy (round-half-even like torch).
    Host-side helper for workload construction only; parity tests take levels from the oracle."""
    b = np.asarray(boxes, np.float64)
    h = b[:, 2] - b[:, 0]
    w = b[:, 3] - b[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        lvl = 4 + np.log2(np.sqrt(np.maximum(h * w, 0)) / (224.0 / np.sqrt(image_hw[0] * image_hw[1])))
    lvl = np.where(np.isfinite(lvl), lvl, -100.0)
    return np.clip(np.rint(lvl), 2, 5).astype(np.int32)


def feature_map(B, C, H, W, seed=1234):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, C, H, W), dtype=np.float32)


def nms_boxes(n, seed=7, kind="rpn", image=1024.0, rounded=False):
    """Config-3 box generator.  kind='rpn': n/20 cluster centres, N(0,8px) jitter, side
    log-uniform [32,512], ratio in {1/2,1,2}; kind='uniform': low-overlap worst case.
    Returns f32[n,4] (y1,x1,y2,x2) in pixels, clipped to [0,image]."""
    rng = np.random.default_rng(seed)
    if kind == "rpn":
        k = max(n // 20, 1)
        centres = rng.uniform(0, image, (k, 2))
        sides = np.exp(rng.uniform(np.log(32.0), np.log(512.0), k))
        ratios = rng.choice([0.5, 1.0, 2.0], k)
        cid = rng.integers(0, k, n)
        c = centres[cid] + rng.normal(0, 8.0, (n, 2))
        s = sides[cid] * np.exp(rng.normal(0, 0.05, n))
        r = ratios[cid]
        h = s * np.sqrt(r)
        w = s / np.sqrt(r)
    else:
        c = rng.uniform(0, image, (n, 2))
        s = np.exp(rng.uniform(np.log(8.0), np.log(64.0), n))
        r = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
        h = s * np.sqrt(r)
        w = s / np.sqrt(r)
    b = np.stack([c[:, 0] - h / 2, c[:, 1] - w / 2, c[:, 0] + h / 2, c[:, 1] + w / 2], 1)
    b = np.clip(b, 0, image).astype(np.float32)
    if rounded:
        b = np.rint(b).astype(np.float32)
    return b


def nms_scores(n, seed=8, ties=False):
    """Tie-free: a random permutation of linspace(0,1,n).  ties=True: quantised to 256 levels."""
    rng = np.random.default_rng(seed)
    s = rng.permutation(np.linspace(0.0, 1.0, n)).astype(np.float32)
    if ties:
        s = (np.floor(s * 255.0) / 255.0).astype(np.float32)
    return s


def label_map(H=1024, W=1024, n=20, seed=2024, min_piece=64):
    """Config-4 painter's-order label map (writer rule: COCOA_D2S_TO_OurFormate.ipynb
    `reLayerMask`): object i is drawn over objects < i ... here object 0 is front-most.
    bit i = object i visible, bit 32+i = object i present but occluded.  Pieces (distinct label
    values) smaller than `min_piece` pixels are zeroed like the writer does.  Returns u64[H,W]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    present = np.zeros((n, H, W), bool)
    for i in range(n):
        cy, cx = rng.uniform(0.15, 0.85) * H, rng.uniform(0.15, 0.85) * W
        ry, rx = rng.uniform(0.04, 0.22) * H, rng.uniform(0.04, 0.22) * W
        if rng.random() < 0.5:
            present[i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        else:
            present[i] = (np.abs(yy - cy) <= ry) & (np.abs(xx - cx) <= rx)
    label = np.zeros((H, W), np.uint64)
    covered = np.zeros((H, W), bool)
    for i in range(n):                       # object 0 front-most
        vis = present[i] & ~covered
        occ = present[i] & covered
        label |= vis.astype(np.uint64) << np.uint64(i)
        label |= occ.astype(np.uint64) << np.uint64(32 + i)
        covered |= present[i]
    vals, inv, counts = np.unique(label, return_inverse=True, return_counts=True)
    small = counts < min_piece
    small[vals == 0] = False
    label = np.where(small[inv].reshape(H, W), np.uint64(0), label)
    return label


def pyramid_anchors(image=1024, scales=(32, 64, 128, 256, 512), ratios=(0.5, 1.0, 2.0), strides=(4, 8, 16, 32, 64)):
    """Anchor pyramid of the reference's configuration (config.py:58-73): for every FPN level a grid of centres at the
    level's stride and one box per aspect ratio -- 261 888 anchors for a 1024^2 image.  f64 [A,4] (y1,x1,y2,x2) pixels."""
    out = []
    for scale, stride in zip(scales, strides):
        n = image // stride
        r = np.asarray(ratios, np.float64)
        h, w = scale / np.sqrt(r), scale * np.sqrt(r)
        cy, cx = np.meshgrid(np.arange(n) * stride, np.arange(n) * stride, indexing="ij")
        cy, cx = cy.reshape(-1, 1).astype(np.float64), cx.reshape(-1, 1).astype(np.float64)
        boxes = np.stack([cy - 0.5 * h, cx - 0.5 * w, cy + 0.5 * h, cx + 0.5 * w], axis=2).reshape(-1, 4)
        out.append(boxes)
    return np.concatenate(out, axis=0)

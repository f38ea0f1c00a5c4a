"""Driver for ncu over the kernels either side of the path: python tools/prof_side.py [unmold|rle] (one warm-up + one call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sln_amodal_b200 import rle, unmold

what = sys.argv[1] if len(sys.argv) > 1 else "unmold"
dev = torch.device("cuda", 0)
n, h, w = 100, 1024, 1024
rng = np.random.default_rng(505)
if what == "rle":
    yy, xx = np.mgrid[0:h, 0:w]
    masks = np.zeros((n, h, w), np.uint8)
    for i in range(n):
        cy, cx = rng.uniform(0.2, 0.8, 2) * h
        ry, rx = rng.uniform(0.05, 0.2, 2) * h
        masks[i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    cols = torch.from_numpy(masks).to(dev).transpose(1, 2).contiguous().view(n, h * w)
    fn = lambda: rle.rle_counts_device(cols)
else:
    mk = torch.from_numpy(rng.random((n, 28, 28)).astype(np.float32)).to(dev)
    bx = np.zeros((n, 4), np.int32)
    for i in range(n):
        bh, bw = rng.integers(40, 400, 2)
        y1, x1 = rng.integers(0, h - bh), rng.integers(0, w - bw)
        bx[i] = (y1, x1, y1 + bh, x1 + bw)
    bxt = torch.from_numpy(bx).to(dev)
    fn = lambda: unmold.unmold_masks(mk, bxt, (h, w))
for _ in range(2):
    fn()
torch.cuda.synchronize()

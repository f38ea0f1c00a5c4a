"""Debug driver for the backward kernels: python tools/dbg_bwd.py C H W N ph pw [B] [exact] -> compares with the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle
from sln_amodal_b200 import ops, synth
C, H, W, N, ph, pw = (int(a) for a in sys.argv[1:7])
B = int(sys.argv[7]) if len(sys.argv) > 7 else 3
modes = [bool(int(sys.argv[8]))] if len(sys.argv) > 8 else [True, False]
rng = np.random.default_rng(7 + C + N)
boxes = synth.roi_boxes(N, seed=11 + N, outside_frac=0.1, degenerate_frac=0.05)
ind = rng.integers(0, B, N).astype(np.int32)
g = rng.standard_normal((N, C, ph, pw), dtype=np.float32)
want = oracle.crop_and_resize_bwd(g, boxes, ind, (B, C, H, W))
dev = torch.device("cuda", 0)
gt = torch.from_numpy(g).to(dev).contiguous(memory_format=torch.channels_last)
for exact in modes:
    got = ops.crop_and_resize_backward(gt, torch.from_numpy(boxes).to(dev), torch.from_numpy(ind).to(dev), (B, C, H, W), exact=exact)
    torch.cuda.synchronize()
    gn = got.contiguous().cpu().numpy()
    err = np.abs(gn - want).max() / max(np.abs(want).max(), 1e-30)
    print("exact" if exact else "default", "bit-equal" if gn.tobytes() == want.tobytes() else "differs", "max rel err %.3e" % err, flush=True)

import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sln_amodal_b200 import ops, synth
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
maps = [torch.randn((2, 256, s, s), device=dev).contiguous(memory_format=torch.channels_last) for s in (32, 16)]
boxes = torch.from_numpy(synth.roi_boxes(96, seed=1, outside_frac=0.1)).to(dev)
ind = torch.from_numpy(rng.integers(0, 2, 96).astype(np.int32)).to(dev)
level = torch.from_numpy(rng.integers(0, 2, 96).astype(np.int32)).to(dev)
for p in (7, 14):
    g = torch.randn((96, 256, p, p), device=dev).contiguous(memory_format=torch.channels_last)
    ops.pyramid_crop_backward(g, boxes, ind, level, [tuple(m.shape) for m in maps])
torch.cuda.synchronize(); print("done")

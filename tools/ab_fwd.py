"""Forward crop timing: python tools/ab_fwd.py"""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sln_amodal_b200 import ops
dev = torch.device("cuda", 0); torch.manual_seed(0)
boxes_np, ind_np, level_np = bench.make_workload()
maps = [torch.randn((8, 256, s, s), device=dev).contiguous(memory_format=torch.channels_last) for s in bench.LEVEL_SIDES]
boxes, box_ind, level = (torch.from_numpy(a).to(dev) for a in (boxes_np, ind_np, level_np))
import hashlib
res = {}
for p in (7, 14, 16):
    for _ in range(3): ops.pyramid_crop_forward(maps, boxes, box_ind, level, p, p, 0.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.pyramid_crop_forward(maps, boxes, box_ind, level, p, p, 0.0)
    b.record(); torch.cuda.synchronize()
    out = ops.pyramid_crop_forward(maps, boxes, box_ind, level, p, p, 0.0)
    res[p] = (round(a.elapsed_time(b) / 10, 4), hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:8])
print("fwd ms", os.environ.get("SLN_FWD_ASYNC", "0"), res)

"""Diagnostic: floor of the backward's write-out path.  Times (a) torch zero_() of the four config-2 gradient maps, (b) the
backward with zero ROIs (tile machinery + zero write-out only), (c) with ROIs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sln_amodal_b200 import ops, synth
dev = torch.device("cuda", 0)
sizes = [(8, 256, s, s) for s in bench.LEVEL_SIDES]
maps = [torch.empty(sz, device=dev).contiguous(memory_format=torch.channels_last) for sz in sizes]
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("zero_() of the maps (0.713 GB): %.4f ms" % ev(lambda: [m.zero_() for m in maps]))
n = 8000
boxes_np = synth.roi_boxes(n, seed=4321)
level_np = (synth.fpn_level(boxes_np) - 2).astype(np.int32)
ind_np = np.repeat(np.arange(8, dtype=np.int32), 1000)
boxes, ind, level = (torch.from_numpy(a).to(dev) for a in (boxes_np, ind_np, level_np))
for p in (7, 14):
    g = torch.randn((n, 256, p, p), device=dev).contiguous(memory_format=torch.channels_last)
    g0 = g[:0]
    print("pool %d: bwd with 0 ROIs %.4f ms, with 8000 ROIs %.4f ms, read of grads alone (sum) %.4f ms" % (
        p, ev(lambda: ops.pyramid_crop_backward(g0, boxes[:0], ind[:0], level[:0], sizes)),
        ev(lambda: ops.pyramid_crop_backward(g, boxes, ind, level, sizes)), ev(lambda: g.sum())))
    # all ROIs far outside: prep + tile machinery, no hits
    far = boxes + 5.0
    print("   ROIs all outside the maps (prep runs, no hits): %.4f ms" % ev(lambda: ops.pyramid_crop_backward(g, far, ind, level, sizes)))

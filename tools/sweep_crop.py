"""Config-2 sweep: RoIAlign fwd / bwd time and fraction of the HBM roofline for 8 images, C=256, P2-P5,
N/img in {1000, 2000, 4000}, pools {7, 14, 16}.   python tools/sweep_crop.py"""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sln_amodal_b200 import ops, synth
dev = torch.device("cuda", 0)
peak, src = bench.measured_peak_gbs()
maps = [torch.randn((8, 256, s, s), device=dev).contiguous(memory_format=torch.channels_last) for s in bench.LEVEL_SIDES]
sizes = [tuple(m.shape) for m in maps]
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
rows = []
for per in (1000, 2000, 4000):
    n = 8 * per
    boxes_np = synth.roi_boxes(n, seed=4321)
    level_np = (synth.fpn_level(boxes_np) - 2).astype(np.int32)
    ind_np = np.repeat(np.arange(8, dtype=np.int32), per)
    boxes, ind, level = (torch.from_numpy(a).to(dev) for a in (boxes_np, ind_np, level_np))
    bench.IMAGES_PER_GPU = 8
    for p in (7, 14, 16):
        g = torch.randn((n, 256, p, p), device=dev).contiguous(memory_format=torch.channels_last)
        tf = t(lambda: ops.pyramid_crop_forward(maps, boxes, ind, level, p, p, 0.0))
        tb = t(lambda: ops.pyramid_crop_backward(g, boxes, ind, level, sizes))
        bf = bench.fwd_bytes(boxes_np, ind_np, level_np, p)
        bb = sum(bench.bwd_bytes(int((level_np == l).sum()), side, p) for l, side in enumerate(bench.LEVEL_SIDES))
        rows.append({"rois_per_img": per, "pool": p, "fwd_ms": round(tf, 4), "fwd_frac": round(bf / tf / 1e6 / peak, 3),
                     "bwd_ms": round(tb, 4), "bwd_frac": round(bb / tb / 1e6 / peak, 3),
                     "fwd_bwd_Mrois_s": round(n / (tf + tb) / 1e3, 2)})
        del g
print(json.dumps({"peak_gbs": peak, "peak_source": src, "rows": rows}))

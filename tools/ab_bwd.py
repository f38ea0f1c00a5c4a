"""A/B timing of crop backward variants: python tools/ab_bwd.py libA.so libB.so ..."""
import os, sys, shutil, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, torch, numpy as np
sys.path.insert(0, %r)
import bench
from sln_amodal_b200 import ops
dev = torch.device("cuda", 0)
import hashlib
res = {}
for per in [int(v) for v in os.environ.get('AB_PER', '1000,4000').split(',')]:
    bench.ROIS_PER_IMAGE = per
    boxes_np, ind_np, level_np = bench.make_workload()
    maps = [torch.randn((8, 256, s, s), device=dev).contiguous(memory_format=torch.channels_last) for s in bench.LEVEL_SIDES]
    boxes, box_ind, level = (torch.from_numpy(a).to(dev) for a in (boxes_np, ind_np, level_np))
    sizes = [tuple(m.shape) for m in maps]
    gen = torch.Generator(device=dev); gen.manual_seed(per)
    for p in (7, 14, 16):
        g = torch.randn((8 * per, 256, p, p), device=dev, generator=gen).contiguous(memory_format=torch.channels_last)
        plan = ops.pyramid_crop_backward_plan(boxes, box_ind, level, sizes, 256, p, p)
        for _ in range(3): out = ops.pyramid_crop_backward(g, boxes, box_ind, level, sizes, plan=plan)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): ops.pyramid_crop_backward(g, boxes, box_ind, level, sizes, plan=plan)
        b.record(); torch.cuda.synchronize()
        ex = ops.pyramid_crop_backward(g, boxes, box_ind, level, sizes, exact=True)
        h = hashlib.sha1(b"".join(o.cpu().numpy().tobytes() for o in out + ex)).hexdigest()[:8]
        res["%%d/%%d" %% (per, p)] = (round(a.elapsed_time(b) / 10, 4), h)
        del g, out, ex
print(res)
''' % ROOT
for lib in sys.argv[1:]:
    dst = os.path.join(ROOT, "sln_amodal_b200", "libsln_b200.so")
    if os.path.abspath(lib) != dst:
        shutil.copy(lib, dst)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip(), out.stderr.strip()[-300:])

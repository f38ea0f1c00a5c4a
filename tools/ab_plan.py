"""A/B of the backward planned ahead (SLN_BWD_PLAN_ONLY on a side stream beside the forward + SLN_BWD_PLANNED) against
planning inside the backward call, on the config-2 workload.  python tools/ab_plan.py [rois_per_img ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from sln_amodal_b200 import ops


def main():
    dev = torch.device("cuda", 0)
    per_img = [int(a) for a in sys.argv[1:]] or [1000, 4000]
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    maps = [torch.randn((bench.IMAGES_PER_GPU, bench.CHANNELS, s, s), device=dev, generator=g).contiguous(memory_format=torch.channels_last)
            for s in bench.LEVEL_SIDES]
    sizes = [tuple(m.shape) for m in maps]
    for per in per_img:
        bench.ROIS_PER_IMAGE = per
        bn, inn, ln = bench.make_workload()
        boxes, ind, level = (torch.from_numpy(a).to(dev) for a in (bn, inn, ln))
        for p in (7, 14, 16):
            gr = torch.randn((bn.shape[0], bench.CHANNELS, p, p), device=dev, generator=g).contiguous(memory_format=torch.channels_last)

            def step(plan_ahead):
                plan = ops.pyramid_crop_backward_plan(boxes, ind, level, sizes, bench.CHANNELS, p, p) if plan_ahead else None
                ops.pyramid_crop_forward(maps, boxes, ind, level, p, p, 0.0)
                return ops.pyramid_crop_backward(gr, boxes, ind, level, sizes, plan=plan)

            a, b = step(True), step(False)
            same = all(torch.equal(x, y) for x, y in zip(a, b))
            del a, b
            row = {}
            for mode in (False, True, False, True):
                for _ in range(3):
                    step(mode)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    step(mode)
                e1.record()
                torch.cuda.synchronize()
                row.setdefault(mode, []).append(round(e0.elapsed_time(e1) / 20, 4))
            print("rois/img %d pool %d fwd+bwd ms: in-backward %s planned-ahead %s identical %s" % (per, p, row[False], row[True], same), flush=True)


if __name__ == "__main__":
    main()

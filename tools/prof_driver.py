"""Short driver for ncu captures: runs one family of kernels a few times on the bench workload.
    python tools/prof_driver.py {crop|nms|proposal|semdist|targets|after} [reps]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sln_amodal_b200 import ops, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "crop"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)

if what == "crop":
    boxes_np, ind_np, level_np = bench.make_workload()
    maps = [torch.randn((bench.IMAGES_PER_GPU, bench.CHANNELS, s, s), device=dev).contiguous(memory_format=torch.channels_last)
            for s in bench.LEVEL_SIDES]
    boxes, box_ind, level = (torch.from_numpy(a).to(dev) for a in (boxes_np, ind_np, level_np))
    for _ in range(reps):
        for p in bench.POOLS:
            g = torch.randn((boxes.shape[0], bench.CHANNELS, p, p), device=dev).contiguous(memory_format=torch.channels_last)
            sizes = [tuple(m.shape) for m in maps]
            # the bench step: ROI lists planned ahead (side stream), forward, backward on the plan
            plan = ops.pyramid_crop_backward_plan(boxes, box_ind, level, sizes, bench.CHANNELS, p, p)
            ops.pyramid_crop_forward(maps, boxes, box_ind, level, p, p, 0.0)
            ops.pyramid_crop_backward(g, boxes, box_ind, level, sizes, plan=plan)
            del g
elif what == "nms":
    for _ in range(reps):
        for n, kind in ((12000, "rpn"), (6000, "rpn"), (12000, "uniform")):
            dets = torch.from_numpy(np.concatenate([synth.nms_boxes(n, seed=7, kind=kind), synth.nms_scores(n, seed=8)[:, None]], 1)).to(dev)
            ops.nms_device(dets, 0.7)
elif what == "proposal":
    A = 261888
    rng = np.random.default_rng(31)
    an = torch.from_numpy(synth.nms_boxes(A, seed=4, kind="rpn")).to(dev)
    fg = rng.permutation(np.linspace(0, 1, A)).astype(np.float32)
    probs = torch.from_numpy(np.stack([1 - fg, fg], 1).astype(np.float32)).to(dev)
    dl = torch.from_numpy((rng.standard_normal((A, 4)) * 0.5).astype(np.float32)).to(dev)
    for _ in range(reps):
        ops.proposal_device(probs, dl, an, 1000, 0.7, (0.1, 0.1, 0.2, 0.2), (1024, 1024))
elif what == "semdist":
    labels = np.stack([synth.label_map(1024, 1024, n=20, seed=2024 + i) for i in range(2)])
    labels = torch.from_numpy(np.tile(labels, (8, 1, 1)).view(np.int64)).to(dev)
    for _ in range(reps):
        planes, n_obj = ops.layer_decode_device(labels, 1, 20)
        ops.edt_sq_device(planes)
elif what == "targets":
    # the section 8(f) rows: head pipeline pieces, detection / RPN targets, plane boxes, COCO RLE
    dev0 = dev
    for _ in range(reps):
        bench.head_pipeline(dev0, cpu=False)
        bench.detection_targets_metric(dev0, cpu=False)
        bench.rpn_targets_metric(dev0, cpu=False)
        bench.rle_metric(dev0, 6650.0, cpu=False)
elif what == "after":
    # the steps either side of the path added last: mask paste (unmold) and the RPN re-layout
    for _ in range(reps):
        bench.unmold_metric(dev, 6650.0, cpu=False)
        bench.rpn_pack_metric(dev, 6650.0, cpu=False)
        bench.resize_image_metric(dev, cpu=False)
torch.cuda.synchronize()
print("done", what)

"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sln_amodal_b200 import ops, synth, sem_dist_targets
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
# crop fwd/bwd (tile kernel for 7x7, strip kernel for 14x14), 2 levels
maps = [torch.randn((2, 256, s, s), device=dev).contiguous(memory_format=torch.channels_last) for s in (32, 16)]
boxes_np = synth.roi_boxes(96, seed=1, outside_frac=0.1)
boxes = torch.from_numpy(boxes_np).to(dev)
ind = torch.from_numpy(rng.integers(0, 2, 96).astype(np.int32)).to(dev)
level = torch.from_numpy(rng.integers(0, 2, 96).astype(np.int32)).to(dev)
for p in (7, 14):
    out = ops.pyramid_crop_forward(maps, boxes, ind, level, p, p, 0.0)
    g = torch.randn_like(out)
    ops.pyramid_crop_backward(g, boxes, ind, level, [tuple(m.shape) for m in maps])
    ops.pyramid_crop_backward(g, boxes, ind, level, [tuple(m.shape) for m in maps], exact=True)
    plan = ops.pyramid_crop_backward_plan(boxes, ind, level, [tuple(m.shape) for m in maps], 256, p, p)     # planned ahead (side stream)
    ops.pyramid_crop_backward(g, boxes, ind, level, [tuple(m.shape) for m in maps], plan=plan)
# NMS: sparse (plain, class-aware), bail-out to dense, dense only
for n, kind in ((700, "rpn"), (3000, "uniform")):
    dets = torch.from_numpy(np.concatenate([synth.nms_boxes(n, seed=2, kind=kind), synth.nms_scores(n, seed=3)[:, None]], 1)).to(dev)
    ops.nms_device(dets, 0.7)
    ops.nms_device(dets, 0.7, sparse_only=True)
    ops.nms_device(dets, 0.7, dense_only=True)
    cls = torch.from_numpy(rng.integers(0, 9, n).astype(np.int32)).to(dev)
    ops.nms_device(dets, 0.3, class_ids=cls)
dup = torch.from_numpy(np.concatenate([(np.array([100, 100, 300, 300], np.float32) + rng.normal(0, 1, (600, 4))).astype(np.float32),
                                       synth.nms_scores(600, seed=4)[:, None]], 1)).to(dev)
ops.nms_device(dup, 0.7)
# proposal layer
A = 20000
an = torch.from_numpy(synth.nms_boxes(A, seed=4, kind="uniform")).to(dev)
fg = rng.permutation(np.linspace(0, 1, A)).astype(np.float32)
probs = torch.from_numpy(np.stack([1 - fg, fg], 1).astype(np.float32)).to(dev)
dl = torch.from_numpy((rng.standard_normal((A, 4)) * 0.5).astype(np.float32)).to(dev)
ops.proposal_device(probs, dl, an, 300, 0.7, (0.1, 0.1, 0.2, 0.2), (1024, 1024), pre_nms_limit=2000)
# layer decode + EDT (packed path needs W >= 128)
lab = synth.label_map(160, 256, n=6, seed=5, min_piece=16)
sem_dist_targets(lab, 3, n_max=6)
# banded EDT kernels: cut last band, single segment, all foreground (runs through every band), noise
rng_e = np.random.default_rng(9)
for shp, p in (((70, 96), 0.97), ((40, 64), 2.0), ((33, 32), 0.5), ((130, 1024), 0.999)):
    ops.edt_sq_device(torch.from_numpy((rng_e.random(shp) < p).astype(np.uint8)[None]).to(dev))
torch.cuda.synchronize()
print("sanitize_small done")

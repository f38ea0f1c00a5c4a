"""A/B of the programmatic dependent launches (SLN_PDL=1 / 0, read by the library on every launch) for the NMS and
proposal-layer launch chains: same process, same inputs, identical results required; plus the unmold kernel's time.
    python tools/ab_pdl.py  ->  one JSON line (also written to gpurun_out/ab_pdl.json)"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sln_amodal_b200 import ops, synth, unmold, rle  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t_us(fn, reps=30):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return round(float(np.median(ts)), 1)


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            fn()
    return g


cases = {}
for n, kind in ((12000, "rpn"), (12000, "uniform"), (6000, "rpn"), (2000, "rpn")):
    d = torch.from_numpy(np.concatenate([synth.nms_boxes(n, seed=7, kind=kind), synth.nms_scores(n, seed=8)[:, None]], 1)).to(dev)
    cases["nms %d %s sparse-only" % (n, kind)] = (lambda d=d: ops.nms_device(d, 0.7, sparse_only=True))
    if n == 12000 and kind == "rpn":
        cases["nms 12000 rpn with dense early-outs"] = (lambda d=d: ops.nms_device(d, 0.7))
rng = np.random.default_rng(81)
d81 = torch.from_numpy(np.concatenate([synth.nms_boxes(12000, seed=9, rounded=True), synth.nms_scores(12000, seed=8)[:, None]], 1)).to(dev)
c81 = torch.from_numpy(rng.integers(1, 81, 12000).astype(np.int32)).to(dev)
cases["nms 12000 K=81 sparse-only"] = lambda: ops.nms_device(d81, 0.3, class_ids=c81, sparse_only=True)
A = 261888
rng = np.random.default_rng(31)
an = torch.from_numpy(synth.nms_boxes(A, seed=4, kind="rpn")).to(dev)
fg = rng.permutation(np.linspace(0, 1, A)).astype(np.float32)
probs = torch.from_numpy(np.stack([1 - fg, fg], 1).astype(np.float32)).to(dev)
dl = torch.from_numpy((rng.standard_normal((A, 4)) * 0.5).astype(np.float32)).to(dev)
cases["proposal_layer 261888"] = lambda: ops.proposal_device(probs, dl, an, 1000, 0.7, (0.1, 0.1, 0.2, 0.2), (1024, 1024))


def flat(r):
    r = r if isinstance(r, (tuple, list)) else (r,)
    return [x.detach().cpu().numpy() for x in r if isinstance(x, torch.Tensor)]


out = {"pdl": {}, "unmold": {}}
for name, fn in cases.items():
    row = {}
    res = {}
    for mode in ("1", "0"):
        os.environ["SLN_PDL"] = mode
        res[mode] = flat(fn())
        row["us_pdl" + mode] = t_us(fn)
        try:
            g = graphed(fn)
            row["us_graph_pdl" + mode] = t_us(g.replay)
            del g
        except Exception as e:  # noqa: BLE001
            row["graph_error_pdl" + mode] = str(e)[:200]
    row["identical"] = all(np.array_equal(a, b) for a, b in zip(res["1"], res["0"])) and len(res["1"]) == len(res["0"])
    out["pdl"][name] = row
os.environ["SLN_PDL"] = "1"

# unmold: 100 detections of a 1024^2 image, 28x28 masks, boxes like the head's detections
rng = np.random.default_rng(3)
N, H, W = 100, 1024, 1024
masks = torch.from_numpy(rng.random((N, 28, 28)).astype(np.float32)).to(dev)
hw = np.exp(rng.uniform(np.log(24), np.log(600), (N, 2)))
y1x1 = rng.uniform(0, 1, (N, 2)) * (1024 - hw)
boxes_np = np.concatenate([y1x1, y1x1 + hw], 1).astype(np.int32)
boxes = torch.from_numpy(boxes_np).to(dev)
us = t_us(lambda: unmold.unmold_masks(masks, boxes, (H, W)))
planes = unmold.unmold_masks(masks, boxes, (H, W))
out["unmold"] = {"n": N, "image": [H, W], "us": us, "gbs_written": round(N * H * W / us / 1e3, 1),
                 "box_area_frac": round(float(((boxes_np[:, 2] - boxes_np[:, 0]) * (boxes_np[:, 3] - boxes_np[:, 1])).sum() / (N * H * W)), 3),
                 "fg_frac": round(float(planes.float().mean().item()), 4),
                 "unmold_plus_rle_encode_ms": round(t_us(lambda: rle.encode(unmold.unmold_masks(masks, boxes, (H, W))), reps=10) / 1e3, 2)}
line = json.dumps(out)
print(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "ab_pdl.json"), "w").write(line + "\n")

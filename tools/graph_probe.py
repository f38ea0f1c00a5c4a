"""Device-only time of the launch-bound paths: capture one call in a CUDA graph, replay it, time the replays.
    python tools/graph_probe.py
"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sln_amodal_b200 import ops, synth
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=20, fl=True):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        if fl: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return round(float(np.median(ts)), 1)

def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out

for n, kind in ((12000, "rpn"), (6000, "rpn"), (1000, "rpn"), (12000, "uniform")):
    dets = torch.from_numpy(np.concatenate([synth.nms_boxes(n, seed=7, kind=kind), synth.nms_scores(n, seed=8)[:, None]], 1)).to(dev)
    eager = timed(lambda: ops.nms_device(dets, 0.7, sparse_only=True))
    want = ops.nms_device(dets, 0.7)
    g, out = graphed(lambda: ops.nms_device(dets, 0.7, sparse_only=True))
    rep = timed(g.replay)
    k = int(out[1].item())
    ok = k == int(want[1].item()) and torch.equal(out[0][:k], want[0][:k])
    g2, out2 = graphed(lambda: ops.nms_device(dets, 0.7))
    rep2 = timed(g2.replay)
    print("nms", n, kind, "eager", eager, "graph", rep, "graph+fallback launches", rep2, "same result", ok)
A = 261888
rng = np.random.default_rng(31)
an = torch.from_numpy(synth.nms_boxes(A, seed=4, kind="rpn")).to(dev)
fg = rng.permutation(np.linspace(0, 1, A)).astype(np.float32)
probs = torch.from_numpy(np.stack([1 - fg, fg], 1).astype(np.float32)).to(dev)
dl = torch.from_numpy((rng.standard_normal((A, 4)) * 0.5).astype(np.float32)).to(dev)
f = lambda: ops.proposal_device(probs, dl, an, 1000, 0.7, (0.1, 0.1, 0.2, 0.2), (1024, 1024))
eager = timed(f)
want = f()
g, out = graphed(f)
rep = timed(g.replay)
print("proposal eager", eager, "graph", rep, "same", all(torch.equal(a, b) for a, b in zip(out, want)))

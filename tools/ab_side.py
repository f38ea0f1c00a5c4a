"""A/B timing of the kernels either side of the path (COCO RLE, unmold): python tools/ab_side.py libA.so libB.so ...
Each library is copied over sln_amodal_b200/libsln_b200.so and timed in its own process; an output checksum is printed so
variants can be compared for identity."""
import os, sys, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, hashlib, torch, numpy as np
sys.path.insert(0, %r)
from sln_amodal_b200 import rle, unmold
dev = torch.device("cuda", 0)
n, h, w = 100, 1024, 1024
rng = np.random.default_rng(505)
yy, xx = np.mgrid[0:h, 0:w]
masks = np.zeros((n, h, w), np.uint8)
for i in range(n):
    cy, cx = rng.uniform(0.2, 0.8, 2) * h
    ry, rx = rng.uniform(0.05, 0.2, 2) * h
    masks[i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
d = torch.from_numpy(masks).to(dev)
cols = d.transpose(1, 2).contiguous().view(n, h * w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def ev_us(fn, reps=10, fl=True):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if fl: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))
def graph_us(fn, reps=10):
    """fn's device time from a graph of `reps` (flush, fn) pairs minus a graph of `reps` flushes: no host launch latency"""
    def cap(with_fn):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); flush.zero_()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    flush.zero_()
                    if with_fn: fn()
        return g
    res = []
    for g in (cap(True), cap(False)):
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3 / reps)
        res.append(float(np.median(ts)))
    return res[0] - res[1]
cn, mm = rle.rle_counts_device(cols)
cn, mm = cn.cpu().numpy(), mm.cpu().numpy()
hs = hashlib.sha1(mm.tobytes() + b"".join(cn[i, :mm[i]].tobytes() for i in range(n))).hexdigest()[:10]
print("rle   flushed %%.1f (min %%.1f)  graph(flushed) %%.1f us  sha %%s" %% (*ev_us(lambda: rle.rle_counts_device(cols)), graph_us(lambda: rle.rle_counts_device(cols)), hs))
N = 100
mk = torch.from_numpy(rng.random((N, 28, 28)).astype(np.float32)).to(dev)
bx = np.zeros((N, 4), np.int32)
for i in range(N):
    bh, bw = rng.integers(40, 400, 2)
    y1, x1 = rng.integers(0, h - bh), rng.integers(0, w - bw)
    bx[i] = (y1, x1, y1 + bh, x1 + bw)
bxt = torch.from_numpy(bx).to(dev)
o = unmold.unmold_masks(mk, bxt, (h, w))
hs = hashlib.sha1(o.cpu().numpy().tobytes()).hexdigest()[:10]
print("unmold flushed %%.1f (min %%.1f)  graph(flushed) %%.1f us  sha %%s" %% (*ev_us(lambda: unmold.unmold_masks(mk, bxt, (h, w))), graph_us(lambda: unmold.unmold_masks(mk, bxt, (h, w))), hs))
''' % ROOT
for lib in sys.argv[1:]:
    dst = os.path.join(ROOT, "sln_amodal_b200", "libsln_b200.so")
    if os.path.abspath(lib) != dst:
        shutil.copy(lib, dst)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print("==", os.path.basename(lib)); print(out.stdout.strip()); print(out.stderr.strip()[-400:])

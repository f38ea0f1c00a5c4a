"""A/B timing of NMS variants: python tools/ab_nms.py libA.so libB.so ..."""
import os, sys, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from sln_amodal_b200 import ops, synth
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=20):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return round(float(np.median(ts)), 1)
out = {}
for n, kind in ((12000, "rpn"), (12000, "uniform"), (6000, "rpn"), (1000, "rpn")):
    dets = torch.from_numpy(np.concatenate([synth.nms_boxes(n, seed=7, kind=kind), synth.nms_scores(n, seed=8)[:, None]], 1)).to(dev)
    out["%%d %%s" %% (n, kind)] = (t(lambda: ops.nms_device(dets, 0.7, sparse_only=True)), t(lambda: ops.nms_device(dets, 0.7)))
rng = np.random.default_rng(81)
dets = torch.from_numpy(np.concatenate([synth.nms_boxes(12000, seed=9, rounded=True), synth.nms_scores(12000, seed=8)[:, None]], 1)).to(dev)
cls = torch.from_numpy(rng.integers(1, 81, 12000).astype(np.int32)).to(dev)
out["12000 K=81"] = t(lambda: ops.nms_device(dets, 0.3, class_ids=cls))
print(out)
''' % ROOT
for lib in sys.argv[1:]:
    dst = os.path.join(ROOT, "sln_amodal_b200", "libsln_b200.so")
    if os.path.abspath(lib) != dst:
        shutil.copy(lib, dst)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip(), out.stderr.strip()[-300:])

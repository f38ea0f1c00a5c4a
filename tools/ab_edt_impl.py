"""EDT on the config-4 workload (320 maps of 1024^2): banded kernels against the round-1 whole-column kernels
(SLN_EDT_IMPL=legacy), same process, CUDA events, identical output checked.  python tools/ab_edt_impl.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sln_amodal_b200 import ops, synth
dev = torch.device("cuda", 0)
labels = np.stack([synth.label_map(1024, 1024, n=20, seed=2024 + i) for i in range(4)])
labels = torch.from_numpy(np.tile(labels, (4, 1, 1)).view(np.int64)).to(dev)
for L in (1, 2):
    planes, n_obj = ops.layer_decode_device(labels, L, 20)
    res = {}
    for impl in ("legacy", "band"):
        os.environ["SLN_EDT_IMPL"] = impl
        for _ in range(3): out = ops.edt_sq_device(planes)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = ops.edt_sq_device(planes); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        res[impl] = out
        nbytes = 5 * planes.numel()
        print(f"L={L} {impl}: median {np.median(ts):.1f} us min {min(ts):.1f} us  {nbytes / np.median(ts) / 1e3:.0f} GB/s  ({planes.numel() >> 20} maps)")
    print("identical:", torch.equal(res["legacy"], res["band"]))

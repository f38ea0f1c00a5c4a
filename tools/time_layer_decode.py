import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from sln_amodal_b200 import ops, synth
dev = torch.device('cuda', 0)
labels = np.stack([synth.label_map(1024, 1024, n=20, seed=2024 + (i % 4)) for i in range(4)])
labels = torch.from_numpy(np.tile(labels, (4, 1, 1)).view(np.int64)).to(dev)
for L in (1, 2):
    fn = lambda: ops.layer_decode_device(labels, L, 20)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    by = 16 * 1024 * 1024 * (8 + 20 * L)
    print("L", L, "us", np.median(ts), "frac", by / np.median(ts) / 1e3 / 6532.9)

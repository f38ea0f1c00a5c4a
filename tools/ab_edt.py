"""A/B timing of EDT variants: python tools/ab_edt.py libA.so libB.so ..."""
import os, sys, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from sln_amodal_b200 import ops, synth
dev = torch.device("cuda", 0)
labels = np.stack([synth.label_map(1024, 1024, n=20, seed=2024 + i) for i in range(4)])
labels = torch.from_numpy(np.tile(labels, (4, 1, 1)).view(np.int64)).to(dev)
planes, n_obj = ops.layer_decode_device(labels, 1, 20)
for _ in range(2): ops.edt_sq_device(planes)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): ops.edt_sq_device(planes)
b.record(); torch.cuda.synchronize()
print("edt 320 maps: %%.1f us" %% (a.elapsed_time(b) / 5 * 1e3))
''' % ROOT
for lib in sys.argv[1:]:
    dst = os.path.join(ROOT, "sln_amodal_b200", "libsln_b200.so")
    if os.path.abspath(lib) != dst:
        shutil.copy(lib, dst)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip(), out.stderr.strip()[-300:])

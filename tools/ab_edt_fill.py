"""A/B of the EDT column pass: fill role inside the envelope launch (SLN_EDT_FILL_CTAS=0) against the split form with F fat
fill CTAs that own their SMs (csrc/semdist.cu), same process, same planes (config 4: 320 maps of 1024^2), identical output
required.    python tools/ab_edt_fill.py [F ...]  ->  one JSON line (also gpurun_out/ab_edt_fill.json)"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sln_amodal_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
labels = np.stack([synth.label_map(1024, 1024, n=20, seed=2024 + i) for i in range(4)])
labels = torch.from_numpy(np.tile(labels, (4, 1, 1)).view(np.int64)).to(dev)
planes, n_obj = ops.layer_decode_device(labels, 1, 20)


def t_us(fn, reps=7):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return round(float(np.median(ts)), 1)


out = {}
os.environ["SLN_EDT_FILL_CTAS"] = "0"
ref = ops.edt_sq_device(planes).clone()
for F in [0] + [int(a) for a in sys.argv[1:]] or [0, 16, 24, 32, 40, 56]:
    os.environ["SLN_EDT_FILL_CTAS"] = str(F)
    row = {"us": t_us(lambda: ops.edt_sq_device(planes))}
    got = ops.edt_sq_device(planes)
    row["identical"] = bool(torch.equal(got, ref))
    del got
    if F:
        os.environ["SLN_PDL"] = "0"
        row["us_without_pdl"] = t_us(lambda: ops.edt_sq_device(planes), reps=3)
        os.environ["SLN_PDL"] = "1"
    out["F=%d" % F] = row
line = json.dumps(out)
print(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "ab_edt_fill.json"), "w").write(line + "\n")

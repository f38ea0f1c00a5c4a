"""Times only the side metrics of bench.py (NMS sweep, proposal layer, layer decode, EDT) on cuda:0.
    python tools/time_side.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

peak, _ = bench.measured_peak_gbs()
print(json.dumps(bench.side_metrics(torch.device("cuda", 0), peak)))

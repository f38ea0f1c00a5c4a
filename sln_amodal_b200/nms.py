"""Drop-in for the reference's nms/nms_wrapper.py and nms/pth_nms.py.

    nms(dets, thresh) -> LongTensor[k]      (nms_wrapper.py:14-17 -> pth_nms.py:5-51)

dets f32 [n,5] rows (y1, x1, y2, x2, score); returns indices into dets of the kept boxes,
score-descending, on dets' device.  Semantics are those of the reference's CPU extension
(nms/src/nms.c: "+1" areas, un-fused fp32, `ovr >= thresh`), which BASELINE.json names as
the parity target; the reference's own GPU branch differs (`>`; and pth_nms.py:32-49 hands
the kernel an unsorted tensor).  Ties in score are visited in index order (stable).
"""
from __future__ import annotations

import torch

from . import ops


def pth_nms(dets, thresh):
    """dets has to be a CUDA tensor."""
    # the sparse pipeline alone; one host read, like the reference's `keep[:num_out[0]]` (pth_nms.py:24), which also
    # tells whether the input was outside the sparse contract (k < 0): then the dense pipeline produces the result
    keep, num = ops.nms_device(dets, thresh, sparse_only=True)
    k = int(num.item())
    if k < 0:
        keep, num = ops.nms_device(dets, thresh, dense_only=True)
        k = int(num.item())
    return keep[:k]


def nms(dets, thresh):
    """Dispatch to the device NMS.  Accept dets as tensor."""
    return pth_nms(dets, thresh)


def batched_nms(boxes, scores, class_ids, thresh, max_keep=0):
    """Per-class NMS in one call (the loop of refine_detections, modal/Functions.py:506-525):
    a box can only be suppressed by a higher-scoring box of the same class.  Returns the kept
    indices, score-descending across all classes."""
    dets = torch.cat((boxes.float(), scores.float().unsqueeze(1)), dim=1)
    keep, num = ops.nms_device(dets, thresh, class_ids=class_ids, max_keep=max_keep, sparse_only=True)
    k = int(num.item())
    if k < 0:
        keep, num = ops.nms_device(dets, thresh, class_ids=class_ids, max_keep=max_keep, dense_only=True)
        k = int(num.item())
    return keep[:k]

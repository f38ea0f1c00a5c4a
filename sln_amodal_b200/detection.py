"""Detection refinement with batched per-class NMS (reference modal/Functions.py:453-575).

refine_detections keeps the reference's signature and return convention; the python loop over
class ids (:506-525 -- one nonzero + sort + nms + unique1d per class, each with host syncs) is
one class-aware NMS call.
"""
from __future__ import annotations

import numpy as np
import torch

from .nms import batched_nms


def apply_box_deltas(boxes, deltas):
    """Functions.py:77-98 (torch ops, same order)."""
    height = boxes[:, 2] - boxes[:, 0]
    width = boxes[:, 3] - boxes[:, 1]
    center_y = boxes[:, 0] + 0.5 * height
    center_x = boxes[:, 1] + 0.5 * width
    center_y = center_y + deltas[:, 0] * height
    center_x = center_x + deltas[:, 1] * width
    height = height * torch.exp(deltas[:, 2])
    width = width * torch.exp(deltas[:, 3])
    y1 = center_y - 0.5 * height
    x1 = center_x - 0.5 * width
    y2 = y1 + height
    x2 = x1 + width
    return torch.stack([y1, x1, y2, x2], dim=1)


def _refine_detections_fused(rois, probs, deltas, window, config):
    """CUDA path with USE_NMS: one decode launch (argmax, delta decode, clip, round, keep filter), one class-aware
    NMS call, one host read.  Filtered-out boxes ride through the NMS with score -inf and a private class and are
    dropped from the tail of its score-ordered output."""
    from . import ops
    height, width = config.IMAGE_SHAPE[:2]
    std_dev = np.reshape(config.RPN_BBOX_STD_DEV, [4])
    min_conf = float(config.DETECTION_MIN_CONFIDENCE or 0.0)
    dets, cls_nms, class_ids, n_excl = ops.refine_decode_device(rois, probs, deltas, std_dev, (height, width), window, min_conf)
    thr = config.DETECTION_NMS_THRESHOLD
    keep, num = ops.nms_device(dets, thr, class_ids=cls_nms, sparse_only=True)
    k, n_ex = (int(v) for v in torch.cat((num, n_excl)).tolist())          # the one host read
    if k < 0:                                                               # outside the sparse NMS contract
        keep, num = ops.nms_device(dets, thr, class_ids=cls_nms, dense_only=True)
        k = int(num.item())
    k -= n_ex
    if k <= 0:
        return [], []
    keep = keep[:k]                                                         # score-descending (:538-546)
    sel = dets[keep]
    result = torch.cat((sel[:, :4], class_ids[keep].unsqueeze(1).float(), sel[:, 4:5]), dim=1)
    return result, keep


def refine_detections(rois, probs, deltas, window, config):
    """rois [N,4] normalised, probs [N,K], deltas [N,K,4], window (y1,x1,y2,x2) pixels
    -> (detections [M,6] (y1,x1,y2,x2,class_id,score), keep indices) or ([], [])."""
    dev = rois.device
    if rois.is_cuda and config.USE_NMS and rois.shape[0] > 0:
        return _refine_detections_fused(rois, probs, deltas, window, config)
    _, class_ids = torch.max(probs, dim=1)
    idx = torch.arange(class_ids.size(0), device=dev)
    class_scores = probs[idx, class_ids]
    deltas_specific = deltas[idx, class_ids]
    std_dev = torch.from_numpy(np.reshape(config.RPN_BBOX_STD_DEV, [1, 4])).float().to(dev)
    refined = apply_box_deltas(rois, deltas_specific * std_dev)                    # :436-450
    height, width = config.IMAGE_SHAPE[:2]
    refined = refined * torch.tensor([height, width, height, width], dtype=torch.float32, device=dev)
    w = [float(v) for v in window]
    refined = torch.stack([refined[:, 0].clamp(w[0], w[2]), refined[:, 1].clamp(w[1], w[3]),
                           refined[:, 2].clamp(w[0], w[2]), refined[:, 3].clamp(w[1], w[3])], dim=1)   # :423-433
    refined = torch.round(refined)                                                 # :485
    keep_bool = class_ids > 0
    if config.USE_NMS:
        if config.DETECTION_MIN_CONFIDENCE:
            keep_bool = keep_bool & (class_scores >= config.DETECTION_MIN_CONFIDENCE)
        keep = torch.nonzero(keep_bool)[:, 0]
        if keep.numel() == 0:
            return [], []
        nms_keep = batched_nms(refined[keep], class_scores[keep], class_ids[keep],
                                    config.DETECTION_NMS_THRESHOLD)               # :506-525 in one call
        keep = keep[nms_keep]
    else:
        keep = torch.nonzero(keep_bool).view(-1)
        if keep.numel() > 100:                                                     # :528-532
            order = torch.sort(class_scores[keep], descending=True, stable=True)[1]
            keep = keep[order[:100]]
    if keep.numel() == 0:
        return [], []
    order = torch.sort(class_scores[keep], descending=True, stable=True)[1]        # :538-546
    keep = keep[order]
    result = torch.cat((refined[keep], class_ids[keep].unsqueeze(1).float(), class_scores[keep].unsqueeze(1)), dim=1)
    return result, keep

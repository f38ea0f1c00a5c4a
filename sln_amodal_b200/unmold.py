"""Mask paste after the path -- utils.unmold_mask (utils.py:447-465) and MaskRCNN.unmold_detections
(model.py:747-806) with the per-detection python loop replaced by one launch.

    unmold_mask(mask, bbox, image_shape)                               -> u8 [H, W]   (the reference's signature)
    unmold_masks(masks [N,h,w], boxes [N,4], image_shape)              -> u8 [N, H, W] on the device
    unmold_detections(detections, mrcnn_mask, image_shape, window)     -> boxes, class_ids, scores, masks [H, W, N]

The reference resizes with scipy.misc.imresize(interp='bilinear') = scipy's bytescale + Pillow's 8-bit resampler;
sln_unmold_masks reproduces both bit for bit (csrc/unmold.cu).  The planes stay on the device so that
rle.encode() -- the next step of the reference's evaluation (amodal_train.py:371-400) -- takes them without a copy.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def unmold_masks(masks, boxes, image_shape):
    """masks f32 [N, h, w] and boxes int [N, 4] (numpy or tensors; moved to the current CUDA device when they are not
    there yet) -> u8 [N, H, W] CUDA tensor."""
    H, W = int(image_shape[0]), int(image_shape[1])
    if not torch.cuda.is_available():
        raise _lib.SlnError("unmold_masks needs a CUDA device (there is no CPU fallback)")
    dev = masks.device if isinstance(masks, torch.Tensor) and masks.is_cuda else torch.device("cuda", torch.cuda.current_device())
    m = torch.as_tensor(masks, dtype=torch.float32).to(dev).contiguous()
    b = torch.as_tensor(np.asarray(boxes) if not isinstance(boxes, torch.Tensor) else boxes).to(dev).to(torch.int32).contiguous()
    if m.dim() != 3 or b.dim() != 2 or b.shape[1] != 4 or b.shape[0] != m.shape[0]:
        raise _lib.SlnError("unmold_masks: masks [N, h, w] and boxes [N, 4] expected")
    n, mh, mw = m.shape
    out = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    if n:
        with torch.cuda.device(dev):
            check(lib().sln_unmold_masks(ptr(m), n, mh, mw, ptr(b), H, W, ptr(out), stream_ptr()), "sln_unmold_masks")
        _lib.count_launches(1)
    return out


def unmold_mask(mask, bbox, image_shape, device=False):
    """utils.py:447-465: one small float mask -> binary full-image mask, numpy u8 [H, W] like the reference's (callers
    np.stack the results, model.py:794-800); device=True keeps the CUDA tensor.  Like the reference, the box must lie
    inside the image with positive height and width (there the paste raises a shape mismatch; here it is an error too)."""
    m = torch.as_tensor(mask, dtype=torch.float32)
    m = m.reshape([s for s in m.shape if s != 1] or [1, 1])               # mask.squeeze()
    if m.dim() != 2:
        raise _lib.SlnError("unmold_mask: a [height, width] mask expected")
    y1, x1, y2, x2 = (int(v) for v in bbox)
    if not (0 <= y1 < y2 <= int(image_shape[0]) and 0 <= x1 < x2 <= int(image_shape[1])):
        raise ValueError("unmold_mask: box %s is not inside the %dx%d image" % ((y1, x1, y2, x2), image_shape[0], image_shape[1]))
    out = unmold_masks(m[None], np.asarray([[y1, x1, y2, x2]], np.int32), image_shape)[0]
    return out if device else out.cpu().numpy()


def unmold_detections(detections, mrcnn_mask, image_shape, window, return_device_planes=False):
    """model.py:747-806.  detections [N, 6] and mrcnn_mask [N, h, w, num_classes] as numpy (the reference's calling
    convention) or tensors.  The box arithmetic is the reference's numpy (float64 scale / shift, astype(int32),
    zero-area filter); the resize / threshold / paste loop is one launch.  Returns boxes, class_ids, scores and
    masks [H, W, N] (numpy, like the reference) or, with return_device_planes, the u8 [N, H, W] CUDA planes."""
    det = detections.detach().cpu().numpy() if isinstance(detections, torch.Tensor) else np.asarray(detections)
    zero_ix = np.where(det[:, 4] == 0)[0]
    N = zero_ix[0] if zero_ix.shape[0] > 0 else det.shape[0]
    boxes = det[:N, :4]
    class_ids = det[:N, 4].astype(np.int32)
    class_ids[class_ids > 0] = 1                                           # model.py:769 (classes collapsed)
    scores = det[:N, 5]
    idx = np.arange(N)
    if isinstance(mrcnn_mask, torch.Tensor):
        masks = mrcnn_mask[torch.as_tensor(idx, device=mrcnn_mask.device), :, :,
                           torch.as_tensor(class_ids.astype(np.int64), device=mrcnn_mask.device)]
    else:
        masks = np.asarray(mrcnn_mask)[idx, :, :, class_ids]
    h_scale = image_shape[0] / (window[2] - window[0])
    w_scale = image_shape[1] / (window[3] - window[1])
    shift = window[:2]
    scales = np.array([h_scale, w_scale, h_scale, w_scale])
    shifts = np.array([shift[0], shift[1], shift[0], shift[1]])
    boxes = np.multiply(boxes - shifts, scales).astype(np.int32)
    # model.py:789-793 drops zero-area boxes (the product test also passes boxes with NEGATIVE height and width); a box
    # that rounding pushed past the image border or turned inside out makes the reference's paste raise -- here such a
    # detection is dropped together with its score and class, so boxes, scores and masks stay consistent
    H_, W_ = int(image_shape[0]), int(image_shape[1])
    keep = np.where(((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) > 0) & (boxes[:, 2] > boxes[:, 0]) &
                    (boxes[:, 0] >= 0) & (boxes[:, 1] >= 0) & (boxes[:, 2] <= H_) & (boxes[:, 3] <= W_))[0]
    if keep.shape[0] != N:
        boxes, class_ids, scores = boxes[keep], class_ids[keep], scores[keep]
        masks = masks[torch.as_tensor(keep, device=masks.device)] if isinstance(masks, torch.Tensor) else masks[keep]
        N = class_ids.shape[0]
    if N == 0:
        empty = np.empty((0,) + tuple(np.shape(mrcnn_mask)[1:3]))          # model.py:804 (the reference's quirk)
        return boxes, class_ids, scores, empty
    planes = unmold_masks(masks, boxes, image_shape)
    if return_device_planes:
        return boxes, class_ids, scores, planes
    return boxes, class_ids, scores, np.ascontiguousarray(planes.permute(1, 2, 0).cpu().numpy())

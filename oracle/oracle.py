"""CPU oracle for the SLN-Amodal detection-head hot path.

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, from
``__graft_entry__.smoke()`` and from ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs -- never from ``sln_amodal_b200`` (the product).

Two layers:

* ``ref_*``  -- the reference's own unmodified C (``oracle/_ref/*.so`` built by
  ``oracle/Makefile`` from ``/root/reference/roialign/roi_align/src/crop_and_resize.c``
  and ``/root/reference/nms/src/nms.c`` against ``oracle/th_shim``).  Available
  wherever the prebuilt ``.so`` files are present (they travel to the GPU box).
* everything else -- our restatement (``oracle/sln_oracle.c`` + numpy/torch-CPU
  below), each function citing the reference file:line it follows.
  ``tests/test_oracle_pin.py`` pins the restatement to ``ref_*`` bit for bit,
  to the reference's Python (golden fixtures made by
  ``tests/golden/make_golden.py``) and, for the EDT only, to scipy
  ("parity unpinned" by the reference: it has no EDT).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORC_SO = os.path.join(_HERE, "_build", "libsln_oracle.so")
_REF_CROP_SO = os.path.join(_HERE, "_ref", "libref_crop.so")
_REF_NMS_SO = os.path.join(_HERE, "_ref", "libref_nms.so")
_REF_MASK_SO = os.path.join(_HERE, "_ref", "libref_mask.so")
_REF_CUDA_SO = os.path.join(_HERE, "_ref", "libref_cuda.so")


def build(quiet: bool = True) -> None:
    """(Re)build the oracle libraries with oracle/Makefile (make is a no-op when fresh)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _lptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


_orc = None


def _lib():
    global _orc
    if _orc is None:
        if not os.path.exists(_ORC_SO):
            build()
        lib = C.CDLL(_ORC_SO)
        lib.orc_crop_and_resize_fwd.restype = C.c_int
        lib.orc_crop_and_resize_bwd.restype = C.c_int
        lib.orc_nms.restype = C.c_int
        lib.orc_layer_decode.restype = C.c_int
        lib.orc_edt_sq.restype = None
        lib.orc_edt_sq_banded.restype = None
        lib.orc_nms_mask_scan.restype = C.c_long
        _orc = lib
    return _orc


# ---------------------------------------------------------------------------
# restatement: crop_and_resize (crop_and_resize.c:6-154, 157-252)
# ---------------------------------------------------------------------------
def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def crop_and_resize_fwd(image, boxes, box_ind, ph, pw, ext=0.0):
    """image f32[B,C,H,W] (NCHW), boxes f32[N,4] (y1,x1,y2,x2 normalised), box_ind i32[N]
    -> crops f32[N,C,ph,pw].  Raises on an out-of-range box_ind (reference: exit(-1))."""
    image, boxes, box_ind = _f32(image), _f32(boxes).reshape(-1, 4), _i32(box_ind)
    B, Cc, H, W = image.shape
    N = boxes.shape[0]
    out = np.empty((N, Cc, ph, pw), np.float32)
    rc = _lib().orc_crop_and_resize_fwd(_fptr(image), B, Cc, H, W, _fptr(boxes), _iptr(box_ind),
                                        N, ph, pw, C.c_float(ext), _fptr(out))
    if rc != 0:
        raise ValueError("box_ind out of range")
    return out


def crop_and_resize_bwd(grads, boxes, box_ind, image_shape):
    """grads f32[N,C,ph,pw] -> grad_image f32[B,C,H,W] (dense, zero-filled)."""
    grads, boxes, box_ind = _f32(grads), _f32(boxes).reshape(-1, 4), _i32(box_ind)
    B, Cc, H, W = image_shape
    N, Cg, ph, pw = grads.shape
    assert Cg == Cc
    out = np.empty((B, Cc, H, W), np.float32)
    rc = _lib().orc_crop_and_resize_bwd(_fptr(grads), _fptr(boxes), _iptr(box_ind), N, Cc, ph, pw,
                                        _fptr(out), B, H, W)
    if rc != 0:
        raise ValueError("box_ind out of range")
    return out


# ---------------------------------------------------------------------------
# restatement: NMS (nms.c:4-69 behind pth_nms.py:10-24)
# ---------------------------------------------------------------------------
def stable_order(scores):
    """The build's sort contract (SURVEY.md section 7 'Sort tie-breaking'): score descending,
    index ascending among ties.  (pth_nms.py:17 uses torch's default sort, which is
    implementation-defined on ties.)"""
    scores = np.asarray(scores, dtype=np.float32)
    return np.argsort(-scores.astype(np.float64), kind="stable").astype(np.int64)


def nms_areas(dets):
    """pth_nms.py:10-16: areas = (x2 - x1 + 1) * (y2 - y1 + 1), fp32, un-fused."""
    d = _f32(dets)
    return ((d[:, 3] - d[:, 1] + np.float32(1)) * (d[:, 2] - d[:, 0] + np.float32(1))).astype(np.float32)


def nms_given_order(dets, order, areas, thresh):
    """cpu_nms (nms.c:4-69) on explicit `order` and `areas`.  Returns kept indices int64[k]."""
    dets = _f32(dets)
    order = np.ascontiguousarray(order, dtype=np.int64)
    areas = _f32(areas)
    n = dets.shape[0]
    keep = np.empty(max(n, 1), np.int64)
    num = np.zeros(1, np.int64)
    _lib().orc_nms(_fptr(dets), dets.shape[1], _lptr(order), _fptr(areas), C.c_int64(n),
                   C.c_float(thresh), _lptr(keep), _lptr(num))
    return keep[: int(num[0])].copy()


def nms(dets, thresh):
    """nms(dets, thresh) (nms_wrapper.py:14-17 -> pth_nms.py:5-24, CPU branch) with the stable
    sort contract.  dets f32[n,5] = (y1,x1,y2,x2,score)."""
    dets = _f32(dets)
    if dets.shape[0] == 0:
        return np.empty(0, np.int64)
    return nms_given_order(dets, stable_order(dets[:, 4]), nms_areas(dets), thresh)


# ---------------------------------------------------------------------------
# restatement: layer codec (Functions.py:1012-1095, amodal_train.py:236-271)
# ---------------------------------------------------------------------------
def layer_decode(label, L, n_max=32):
    """Closed-form decode (oracle/sln_oracle.c:orc_layer_decode).
    label u64[H,W] -> (u8[n_max,L,H,W], n_obj)."""
    label = np.ascontiguousarray(label, dtype=np.uint64)
    H, W = label.shape
    out = np.empty((n_max, L, H, W), np.uint8)
    n_obj = _lib().orc_layer_decode(label.ctypes.data_as(C.POINTER(C.c_uint64)), H, W, L, n_max,
                                    out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out, int(n_obj)


def layer_decode_loops(label, num_classes):
    """Loop-for-loop restatement of AmodalDataset.load_layer2 (amodal_train.py:236-271) and the
    codec helpers it drives (Functions.py:1012-1095).  Slow (one full-image compare per label
    piece); small cases only.  Returns bool[H,W,L,n_obj] like the reference, or None when the
    reference would fall through to the empty-mask path (n_obj == 0)."""
    label = np.asarray(label, dtype=np.uint64)
    H, W = label.shape
    L = num_classes - 1
    ids = np.unique(label)                       # get_image_labals, Functions.py:1012-1016
    if ids.size and ids[0] == 0:
        ids = ids[1:]
    lo = ids & np.uint64(0xFFFFFFFF)             # max_objectID, Functions.py:1074-1079
    n_obj = 0
    while np.any((lo >> np.uint64(n_obj)) == 1):
        n_obj += 1
    planes = []
    for i in range(n_obj):
        m = np.zeros((H, W, L), bool)
        for v in ids:                            # objectID_to_masks, Functions.py:1020-1033
            if (int(v) >> i) & 1:
                m[..., 0] |= (label == v)        # amodal_train.py:249-250
        for v in ids:
            if (int(v) >> (i + 32)) & 1:
                hi = int(v) >> 32                # maskID_to_objectIDs, Functions.py:1084-1095
                bits = [k for k in range(32) if (hi >> k) & 1]   # number_to_index :1050-1060
                d = bits.index(i) + 1            # objIDs_to_sindistanceLayer :1063-1064 (+1 at amodal_train.py:255)
                if d >= num_classes - 1 - 1:     # amodal_train.py:256-259
                    m[..., -1] |= (label == v)
                else:
                    m[..., d] |= (label == v)
        planes.append(m)
    if not planes:
        return None
    return np.stack(planes, axis=3)


# ---------------------------------------------------------------------------
# restatement: exact squared EDT (not in the reference; pinned to scipy)
# ---------------------------------------------------------------------------
def edt_sq(mask):
    """mask u8[H,W] -> i32[H,W] squared distance to the nearest zero pixel; (H+W)^2 if none."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    H, W = mask.shape
    out = np.empty((H, W), np.int32)
    _lib().orc_edt_sq(mask.ctypes.data_as(C.POINTER(C.c_ubyte)), H, W,
                      out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def edt_sq_banded(mask, band=32):
    """Second, independent restatement of the same transform: row pass + one lower envelope per band of `band` rows,
    every pixel the minimum over the bands in reach (oracle/sln_oracle.c: orc_edt_sq_banded) -- the executable form of
    the banded column pass planned for the CUDA kernel.  Must equal edt_sq() for every band height."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    H, W = mask.shape
    out = np.empty((H, W), np.int32)
    _lib().orc_edt_sq_banded(mask.ctypes.data_as(C.POINTER(C.c_ubyte)), H, W, int(band),
                             out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


# ---------------------------------------------------------------------------
# restatement: proposal_layer / pyramid level routing (torch-CPU arithmetic)
# ---------------------------------------------------------------------------
def apply_box_deltas(boxes, deltas):
    """Functions.py:77-98 in numpy fp32, one rounding per operation (exp from torch CPU so the
    transcendental is the reference's)."""
    import torch
    b = _f32(boxes)
    d = _f32(deltas)
    f = np.float32
    height = b[:, 2] - b[:, 0]
    width = b[:, 3] - b[:, 1]
    cy = b[:, 0] + f(0.5) * height
    cx = b[:, 1] + f(0.5) * width
    cy = cy + d[:, 0] * height
    cx = cx + d[:, 1] * width
    height = height * torch.exp(torch.from_numpy(d[:, 2].copy())).numpy()
    width = width * torch.exp(torch.from_numpy(d[:, 3].copy())).numpy()
    y1 = cy - f(0.5) * height
    x1 = cx - f(0.5) * width
    y2 = y1 + height
    x2 = x1 + width
    return np.stack([y1, x1, y2, x2], axis=1).astype(np.float32)


def clip_boxes(boxes, window):
    """Functions.py:101-111."""
    b = _f32(boxes).copy()
    w = [np.float32(v) for v in window]
    b[:, 0] = np.clip(b[:, 0], w[0], w[2])
    b[:, 1] = np.clip(b[:, 1], w[1], w[3])
    b[:, 2] = np.clip(b[:, 2], w[0], w[2])
    b[:, 3] = np.clip(b[:, 3], w[1], w[3])
    return b


def proposal_layer(rpn_probs, rpn_bbox, anchors, proposal_count, nms_threshold,
                   image_hw=(1024, 1024), std_dev=(0.1, 0.1, 0.2, 0.2), pre_nms_limit=6000,
                   return_aux=False):
    """Functions.py:114-178 for one image.  rpn_probs f32[A,2], rpn_bbox f32[A,4], anchors f32[A,4]
    (pixels) -> normalised boxes f32[k,4], k <= proposal_count (no padding)."""
    probs, bbox, anchors = _f32(rpn_probs), _f32(rpn_bbox), _f32(anchors)
    scores = probs[:, 1]
    deltas = bbox * np.asarray(std_dev, np.float32).reshape(1, 4)      # :137-140
    limit = min(pre_nms_limit, anchors.shape[0])                        # :144
    order = stable_order(scores)[:limit]                                # :145-146 (stable contract)
    top_scores = scores[order]
    boxes = apply_box_deltas(anchors[order], deltas[order])             # :153
    H, W = image_hw
    boxes = clip_boxes(boxes, (0, 0, H, W))                             # :156-158
    dets = np.concatenate([boxes, top_scores[:, None]], axis=1)
    keep = nms(dets, nms_threshold)[:proposal_count]                    # :165-166
    out = boxes[keep] / np.asarray([H, W, H, W], np.float32)            # :170-173
    if return_aux:
        return out.astype(np.float32), dict(order=order, boxes=boxes, keep=keep)
    return out.astype(np.float32)


def roi_levels(boxes, image_hw=(1024, 1024)):
    """modals.py:53-64 evaluated with torch CPU ops (log/sqrt/round are the reference's)."""
    import torch
    b = torch.from_numpy(_f32(boxes).reshape(-1, 4))
    y1, x1, y2, x2 = b.chunk(4, dim=1)
    h = y2 - y1
    w = x2 - x1
    image_area = torch.FloatTensor([float(image_hw[0] * image_hw[1])])
    ln2 = torch.log(torch.FloatTensor([2.0]))
    lvl = 4 + torch.log(torch.sqrt(h * w) / (224.0 / torch.sqrt(image_area))) / ln2
    lvl = lvl.round().int().clamp(2, 5)
    return lvl.view(-1).numpy().astype(np.int32)


def pyramid_roi_align(boxes, feature_maps, pool, image_hw=(1024, 1024), levels=None):
    """modals.py:20-110 for one image: boxes f32[N,4] normalised, feature_maps = [P2..P5] each
    f32[1,C,H_l,W_l] -> f32[N,C,pool,pool] in the original ROI order."""
    boxes = _f32(boxes).reshape(-1, 4)
    if levels is None:
        levels = roi_levels(boxes, image_hw)
    Cc = feature_maps[0].shape[1]
    out = np.zeros((boxes.shape[0], Cc, pool, pool), np.float32)
    for i, lvl in enumerate(range(2, 6)):
        ix = np.nonzero(levels == lvl)[0]
        if ix.size == 0:
            continue
        out[ix] = crop_and_resize_fwd(feature_maps[i], boxes[ix], np.zeros(ix.size, np.int32), pool, pool, 0.0)
    return out


def per_class_nms(boxes, class_ids, scores, thresh):
    """The per-class loop of refine_detections (Functions.py:506-525): for every class id run
    nms() on that class's detections (sorted by score) and union the kept indices.  Returns the
    sorted union of kept indices (intersect1d/unique1d give ascending order, :524-525)."""
    boxes = _f32(boxes).reshape(-1, 4)
    class_ids = np.asarray(class_ids)
    scores = _f32(scores)
    kept = []
    for c in np.unique(class_ids):
        ixs = np.nonzero(class_ids == c)[0]
        order = stable_order(scores[ixs])
        dets = np.concatenate([boxes[ixs][order], scores[ixs][order][:, None]], axis=1)
        k = nms(dets, thresh)
        kept.append(ixs[order[k]])
    if not kept:
        return np.empty(0, np.int64)
    return np.unique(np.concatenate(kept)).astype(np.int64)


def refine_detections(rois, probs, deltas, window, image_hw=(1024, 1024), std_dev=(0.1, 0.1, 0.2, 0.2),
                      use_nms=True, min_confidence=0.3, nms_threshold=0.3):
    """refine_detections (Functions.py:453-557) for one image in numpy fp32 (exp from torch CPU).
    Returns (detections [M,6] = (y1,x1,y2,x2,class_id,score), keep indices [M]); empty arrays when
    nothing survives."""
    rois, probs, deltas = _f32(rois), _f32(probs), _f32(deltas)
    n = rois.shape[0]
    class_ids = probs.argmax(1)                                           # :468
    idx = np.arange(n)
    class_scores = probs[idx, class_ids]
    d = deltas[idx, class_ids] * np.asarray(std_dev, np.float32).reshape(1, 4)   # coordinate_convert :436-450
    refined = apply_box_deltas(rois, d)
    H, W = image_hw
    refined = refined * np.asarray([H, W, H, W], np.float32)
    refined = clip_boxes(refined, window)                                 # clip_to_window :423-433
    refined = np.rint(refined).astype(np.float32)                         # torch.round: half to even, :485
    keep_bool = class_ids > 0
    if use_nms:
        if min_confidence:
            keep_bool = keep_bool & (class_scores >= np.float32(min_confidence))
        keep = np.nonzero(keep_bool)[0]
        if keep.size == 0:
            return np.zeros((0, 6), np.float32), np.zeros(0, np.int64)
        nms_keep = per_class_nms(refined[keep], class_ids[keep], class_scores[keep], nms_threshold)   # :506-525
        keep = keep[nms_keep]
    else:
        keep = np.nonzero(keep_bool)[0]
        if keep.size > 100:                                               # :528-532
            keep = keep[stable_order(class_scores[keep])[:100]]
    if keep.size == 0:
        return np.zeros((0, 6), np.float32), np.zeros(0, np.int64)
    keep = keep[stable_order(class_scores[keep])]                         # :538-546
    det = np.concatenate([refined[keep], class_ids[keep, None].astype(np.float32), class_scores[keep, None]], 1)
    return det.astype(np.float32), keep.astype(np.int64)


# ---------------------------------------------------------------------------
# detection targets (modal/Functions.py:184-416, utils.py:96-117)
# ---------------------------------------------------------------------------
def bbox_overlaps(boxes1, boxes2):
    """Functions.py:184-218: IoU matrix f32[N1,N2]; no +1 convention, one fp32 rounding per op, 0/0 -> NaN."""
    b1, b2 = _f32(boxes1).reshape(-1, 4), _f32(boxes2).reshape(-1, 4)
    y1 = np.maximum(b1[:, None, 0], b2[None, :, 0])
    x1 = np.maximum(b1[:, None, 1], b2[None, :, 1])
    y2 = np.minimum(b1[:, None, 2], b2[None, :, 2])
    x2 = np.minimum(b1[:, None, 3], b2[None, :, 3])
    zero = np.float32(0)
    inter = np.maximum(x2 - x1, zero) * np.maximum(y2 - y1, zero)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    union = a1[:, None] + a2[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(np.float32)


def box_refinement(box, gt_box):
    """utils.py:96-117 (log from torch CPU)."""
    import torch
    box, gt = _f32(box).reshape(-1, 4), _f32(gt_box).reshape(-1, 4)
    half = np.float32(0.5)
    h, w = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1]
    cy, cx = box[:, 0] + half * h, box[:, 1] + half * w
    gh, gw = gt[:, 2] - gt[:, 0], gt[:, 3] - gt[:, 1]
    gcy, gcx = gt[:, 0] + half * gh, gt[:, 1] + half * gw
    with np.errstate(divide="ignore", invalid="ignore"):
        dy, dx = (gcy - cy) / h, (gcx - cx) / w
        dh = torch.log(torch.from_numpy(gh / h)).numpy()
        dw = torch.log(torch.from_numpy(gw / w)).numpy()
    return np.stack([dy, dx, dh, dw], 1).astype(np.float32)


def detection_target_layer(proposals, gt_class_ids, gt_boxes, gt_masks, train_rois=100, positive_ratio=0.7,
                           bbox_std_dev=(0.1, 0.1, 0.2, 0.2), mask_shape=(32, 32)):
    """Functions.py:223-416 for one image, USE_MINI_MASK off.  proposals f32[N,4], gt_class_ids i32[G],
    gt_boxes f32[G,4], gt_masks u8[L,G,H,W] -> (rois f32[R,4], class_ids i32[R], deltas f32[R,4],
    masks f32[R,L,mh,mw]); empty arrays when nothing is sampled.  Sampling draws torch.randperm from the CPU
    generator exactly where the reference does (:288, :361)."""
    import torch
    props = _f32(proposals).reshape(-1, 4)
    ids = np.asarray(gt_class_ids).astype(np.int32).reshape(-1)
    gtb = _f32(gt_boxes).reshape(-1, 4)
    masks = np.asarray(gt_masks)
    n = props.shape[0]
    if (ids < 0).any():                                                   # COCO crowds :253-267
        crowd = np.nonzero(ids < 0)[0]
        keep = np.nonzero(ids > 0)[0]
        crowd_boxes = gtb[crowd]
        ids, gtb, masks = ids[keep], gtb[keep], masks[:, keep]
        no_crowd = np.nanmax(np.where(np.isnan(bbox_overlaps(props, crowd_boxes)), np.inf, bbox_overlaps(props, crowd_boxes)), 1) < np.float32(0.001)
    else:
        no_crowd = np.ones(n, bool)
    ov = bbox_overlaps(props, gtb)                                        # :274
    ovm = np.where(np.isnan(ov), np.inf, ov)                              # torch.max propagates NaN as the maximum
    iou_max = ovm.max(1)
    pos_bool = iou_max >= np.float32(0.5)                                 # NaN >= 0.5 is False in torch; inf stands for NaN here
    pos_bool &= ~np.isnan(ov).any(1)
    L = masks.shape[0]
    mh, mw = mask_shape
    positive_count = 0
    if pos_bool.any():                                                    # :284-346
        pos_idx = np.nonzero(pos_bool)[0]
        want = int(train_rois * positive_ratio)
        perm = torch.randperm(pos_idx.size).numpy()[:want]
        pos_idx = pos_idx[perm]
        positive_count = pos_idx.size
        pos_rois = props[pos_idx]
        assign = ovm[pos_idx].argmax(1)                                   # first maximum
        roi_gt = gtb[assign]
        roi_cls = ids[assign]
        deltas = box_refinement(pos_rois, roi_gt) / np.asarray(bbox_std_dev, np.float32).reshape(1, 4)
        roi_masks = masks[:, assign]                                      # [L,P,H,W]
        box_ids = np.arange(positive_count, dtype=np.int32)
        tm = np.stack([crop_and_resize_fwd(roi_masks[i][:, None].astype(np.float32), pos_rois, box_ids, mh, mw, 0.0)
                       for i in range(L)], 1)                             # [P,L,1,mh,mw]
        tmasks = np.rint(tm[:, :, 0]).astype(np.float32)                  # torch.round: half to even
    neg_bool = (iou_max < np.float32(0.5)) & ~np.isnan(ov).any(1) & no_crowd   # :351-352
    negative_count = 0
    if positive_count > 0 and neg_bool.any():                             # :354-364 (`.size()` is always truthy)
        neg_idx = np.nonzero(neg_bool)[0]
        r = 1.0 / positive_ratio
        negative_count = int(r * positive_count - positive_count)
        perm = torch.randperm(neg_idx.size).numpy()[:negative_count]
        neg_idx = neg_idx[perm]
        negative_count = neg_idx.size
        neg_rois = props[neg_idx]
    if positive_count > 0 and negative_count > 0:                         # :370-384
        rois = np.concatenate([pos_rois, neg_rois], 0)
        cls = np.concatenate([roi_cls, np.zeros(negative_count, np.int32)])
        deltas = np.concatenate([deltas, np.zeros((negative_count, 4), np.float32)], 0)
        tmasks = np.concatenate([tmasks, np.zeros((negative_count, L, mh, mw), np.float32)], 0)
        return rois, cls, deltas.astype(np.float32), tmasks
    if positive_count > 0:
        return pos_rois, roi_cls, deltas.astype(np.float32), tmasks
    return (np.zeros((0, 4), np.float32), np.zeros(0, np.int32), np.zeros((0, 4), np.float32),
            np.zeros((0, L, mh, mw), np.float32))


# ---------------------------------------------------------------------------
# build_rpn_targets (modal/Functions.py:739-847, utils.compute_overlaps utils.py:54-93)
# ---------------------------------------------------------------------------
def compute_overlaps(boxes1, boxes2):
    """utils.py:78-93 in float64 (what numpy computes for float anchors against int32 GT boxes): IoU f64[N1,N2]."""
    b1 = np.asarray(boxes1, np.float64).reshape(-1, 4)
    b2 = np.asarray(boxes2, np.float64).reshape(-1, 4)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    y1 = np.maximum(b2[None, :, 0], b1[:, None, 0])
    y2 = np.minimum(b2[None, :, 2], b1[:, None, 2])
    x1 = np.maximum(b2[None, :, 1], b1[:, None, 1])
    x2 = np.minimum(b2[None, :, 3], b1[:, None, 3])
    inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0)
    union = area2[None, :] + area1[:, None] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / union


def build_rpn_targets(anchors, gt_class_ids, gt_boxes, anchors_per_image=256, std_dev=(0.1, 0.1, 0.2, 0.2), reductions=None):
    """Functions.py:739-847.  Returns (rpn_match i32[A], rpn_bbox f64[anchors_per_image,4]).  Subsampling draws
    np.random.choice from the global numpy generator exactly where the reference does (:806, :814).
    `reductions` = (anchor_iou_max, anchor_iou_argmax, gt_iou_argmax, crowd_iou_max | None) lets a caller supply the
    overlap reductions computed elsewhere (the CUDA kernels); the rest of the function is shared."""
    anchors = np.asarray(anchors)
    ids = np.asarray(gt_class_ids)
    gtb = np.asarray(gt_boxes)
    A = anchors.shape[0]
    rpn_match = np.zeros([A], dtype=np.int32)
    rpn_bbox = np.zeros((anchors_per_image, 4))
    crowd_ix = np.where(ids < 0)[0]
    crowd_boxes = None
    if crowd_ix.shape[0] > 0:
        non_crowd_ix = np.where(ids > 0)[0]
        crowd_boxes = gtb[crowd_ix]
        ids, gtb = ids[non_crowd_ix], gtb[non_crowd_ix]
    if reductions is None:
        ov = compute_overlaps(anchors, gtb)
        a_arg = np.argmax(ov, axis=1)
        a_max = ov[np.arange(A), a_arg]
        g_arg = np.argmax(ov, axis=0)
        c_max = np.amax(compute_overlaps(anchors, crowd_boxes), axis=1) if crowd_boxes is not None else None
    else:
        a_max, a_arg, g_arg, c_max = reductions
    no_crowd = (c_max < 0.001) if c_max is not None else np.ones([A], dtype=bool)
    rpn_match[(a_max < 0.3) & no_crowd] = -1
    rpn_match[g_arg] = 1
    rpn_match[a_max >= 0.7] = 1
    pos = np.where(rpn_match == 1)[0]
    extra = len(pos) - (anchors_per_image // 2)
    if extra > 0:
        rpn_match[np.random.choice(pos, extra, replace=False)] = 0
    neg = np.where(rpn_match == -1)[0]
    extra = len(neg) - (anchors_per_image - np.sum(rpn_match == 1))
    if extra > 0:
        rpn_match[np.random.choice(neg, extra, replace=False)] = 0
    pos = np.where(rpn_match == 1)[0]                                     # :820-845, all positives at once
    if pos.size:
        a, g = anchors[pos], gtb[a_arg[pos]]
        g_h, g_w = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
        a_h, a_w = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
        dy = ((g[:, 0] + 0.5 * g_h) - (a[:, 0] + 0.5 * a_h)) / a_h
        dx = ((g[:, 1] + 0.5 * g_w) - (a[:, 1] + 0.5 * a_w)) / a_w
        rpn_bbox[: pos.size] = np.stack([dy, dx, np.log(g_h / a_h), np.log(g_w / a_w)], 1) / np.asarray(std_dev, np.float64)
    return rpn_match, rpn_bbox


# ---------------------------------------------------------------------------
# the reference's own C, unmodified (oracle/_ref)
# ---------------------------------------------------------------------------
class _TH(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_long * 4), ("nd", C.c_int)]


def _th(a):
    t = _TH()
    t.data = a.ctypes.data
    for i in range(4):
        t.size[i] = a.shape[i] if i < a.ndim else 1
    t.nd = a.ndim
    return t


_ref_crop = None
_ref_nms = None


def ref_available() -> bool:
    return os.path.exists(_REF_CROP_SO) and os.path.exists(_REF_NMS_SO)


def _ref_crop_lib():
    global _ref_crop
    if _ref_crop is None:
        _ref_crop = C.CDLL(_REF_CROP_SO)
        _ref_crop.crop_and_resize_forward.restype = None
        _ref_crop.crop_and_resize_backward.restype = None
    return _ref_crop


def _ref_nms_lib():
    global _ref_nms
    if _ref_nms is None:
        _ref_nms = C.CDLL(_REF_NMS_SO)
        _ref_nms.cpu_nms.restype = C.c_int
    return _ref_nms


def ref_crop_and_resize_fwd(image, boxes, box_ind, ph, pw, ext=0.0):
    """crop_and_resize_forward (crop_and_resize.c:115-154), the reference's own object code.
    box_ind must be in range: the reference exit(-1)s otherwise."""
    image, boxes, box_ind = _f32(image), _f32(boxes).reshape(-1, 4), _i32(box_ind)
    assert box_ind.size == 0 or (box_ind.min() >= 0 and box_ind.max() < image.shape[0])
    out = np.empty((boxes.shape[0], image.shape[1], ph, pw), np.float32)
    ti, tb, tx, to = _th(image), _th(boxes), _th(box_ind), _th(out)
    _ref_crop_lib().crop_and_resize_forward(C.byref(ti), C.byref(tb), C.byref(tx), C.c_float(ext),
                                            C.c_int(ph), C.c_int(pw), C.byref(to))
    return out


def ref_crop_and_resize_bwd(grads, boxes, box_ind, image_shape):
    """crop_and_resize_backward (crop_and_resize.c:157-252), the reference's own object code."""
    grads, boxes, box_ind = _f32(grads), _f32(boxes).reshape(-1, 4), _i32(box_ind)
    assert box_ind.size == 0 or (box_ind.min() >= 0 and box_ind.max() < image_shape[0])
    out = np.empty(tuple(image_shape), np.float32)
    tg, tb, tx, to = _th(grads), _th(boxes), _th(box_ind), _th(out)
    _ref_crop_lib().crop_and_resize_backward(C.byref(tg), C.byref(tb), C.byref(tx), C.byref(to))
    return out


def ref_nms_given_order(dets, order, areas, thresh):
    """cpu_nms (nms.c:4-69), the reference's own object code."""
    dets = _f32(dets)
    order = np.ascontiguousarray(order, dtype=np.int64)
    areas = _f32(areas)
    n = dets.shape[0]
    keep = np.empty(max(n, 1), np.int64)
    num = np.zeros(1, np.int64)
    tk, tn, td, to, ta = _th(keep), _th(num), _th(dets), _th(order), _th(areas)
    _ref_nms_lib().cpu_nms(C.byref(tk), C.byref(tn), C.byref(td), C.byref(to), C.byref(ta),
                           C.c_float(thresh))
    return keep[: int(num[0])].copy()


def ref_nms(dets, thresh):
    dets = _f32(dets)
    if dets.shape[0] == 0:
        return np.empty(0, np.int64)
    return ref_nms_given_order(dets, stable_order(dets[:, 4]), nms_areas(dets), thresh)


# ---------------------------------------------------------------------------
# COCO run-length codec (cocoapi/common/maskApi.c:32-47, 204-216)
# ---------------------------------------------------------------------------
def rle_encode(mask_flat):
    """maskApi.c:32-41 for one mask given in the memory order to encode (pycocotools passes column-major planes):
    u32 counts, alternating runs starting with the run of zeros (possibly empty)."""
    t = np.asarray(mask_flat, np.uint8).reshape(-1)
    prev = np.concatenate([np.zeros(1, np.uint8), t[:-1]])
    pos = np.nonzero(t != prev)[0]
    edges = np.concatenate([[0], pos, [t.size]])
    return np.diff(edges).astype(np.uint32)


def rle_to_string(counts):
    """maskApi.c:204-216: LEB128-like, 6 bits per char, ascii 48-111, counts[i] delta-coded against counts[i-2] for i > 2."""
    out = bytearray()
    c = [int(v) for v in counts]
    for i, v in enumerate(c):
        x = v - c[i - 2] if i > 2 else v
        more = True
        while more:
            ch = x & 0x1f
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


class _RLE(C.Structure):
    _fields_ = [("h", C.c_ulong), ("w", C.c_ulong), ("m", C.c_ulong), ("cnts", C.POINTER(C.c_uint))]


_ref_mask = None


def _ref_mask_lib():
    global _ref_mask
    if _ref_mask is None:
        lib = C.CDLL(_REF_MASK_SO)
        lib.rleToString.restype = C.c_void_p
        _ref_mask = lib
    return _ref_mask


def ref_mask_available() -> bool:
    return os.path.exists(_REF_MASK_SO)


def ref_rle_encode(masks, h, w):
    """The reference's own rleEncode + rleToString on n masks [n, h*w] (memory order as given).
    Returns [(counts u32[m], string bytes)]."""
    lib = _ref_mask_lib()
    m = np.ascontiguousarray(masks, np.uint8).reshape(-1, h * w)
    n = m.shape[0]
    R = (_RLE * n)()
    lib.rleEncode(R, m.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_ulong(h), C.c_ulong(w), C.c_ulong(n))
    out = []
    libc = C.CDLL(None)
    for i in range(n):
        cnt = np.ctypeslib.as_array(R[i].cnts, shape=(R[i].m,)).copy()
        sp = lib.rleToString(C.byref(R[i]))
        out.append((cnt, C.string_at(sp)))
        libc.free(C.c_void_p(sp))
        lib.rleFree(C.byref(R[i]))
    return out


# ---------------------------------------------------------------------------
# unmold_mask (utils.py:447-465) -- SURVEY.md section 8(f), row 3: the step after the path.
#
# The reference calls scipy.misc.imresize(mask, (y2-y1, x2-x1), interp='bilinear'), i.e. two third-party pieces
# that are NOT in /root/reference:
#   * scipy.misc.pilutil.bytescale / toimage / imresize (scipy 1.0-1.2; removed in scipy 1.3, so absent here):
#     restated below from its published source, with the float32 scalar promotion of the numpy 1.x it ran on;
#   * PIL.Image.resize(size, resample=BILINEAR) = Pillow's libImaging/Resample.c (8-bit, 22-bit fixed-point
#     coefficients, horizontal pass first, 8-bit intermediate).  Pillow IS installed here (12.2.0), so
#     unmold_mask_pil() below runs the real library and pil_resize_bilinear_u8() -- the restatement the CUDA
#     kernel follows -- is pinned to it bit for bit by tests/test_oracle_pin.py.
# ---------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2


def bytescale_f32(data):
    """scipy.misc.bytescale(data) with cmin / cmax = data.min() / data.max(), high=255, low=0, for float32 input
    under numpy-1.x promotion: scale = f32(255.0 / f64(cmax - cmin)); ((data - cmin) * scale).clip(0, 255) + 0.5 in
    float32, truncated to uint8."""
    d = np.ascontiguousarray(data, dtype=np.float32)
    cmin, cmax = d.min(), d.max()
    cscale = np.float32(cmax - cmin)
    if cscale == 0:
        cscale = np.float32(1)
    scale = np.float32(255.0 / float(cscale))
    b = (d - cmin).astype(np.float32) * scale
    b = np.clip(b, np.float32(0), np.float32(255)).astype(np.float32) + np.float32(0.5)
    return b.astype(np.float32).astype(np.uint8)


def pil_bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0), box = whole axis.
    -> (xmin i32 [out], count i32 [out], coeff i32 [out, ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C cast: truncation towards zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(max(xmax, 0), np.float64)
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0.0 else v
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(0.5 + k * float(1 << PIL_PRECISION_BITS))
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, kk


def _pil_pass(src, xmin, cnt, kk):
    """one 8bpc pass along the last axis: out[r, xx] = clip8((2^21 + sum_x src[r, xmin+x] * k[x]) >> 22)"""
    out = np.zeros((src.shape[0], xmin.shape[0]), np.uint8)
    s = src.astype(np.int64)
    for xx in range(xmin.shape[0]):
        acc = np.full(src.shape[0], 1 << (PIL_PRECISION_BITS - 1), np.int64)
        for x in range(cnt[xx]):
            acc += s[:, xmin[xx] + x] * int(kk[xx, x])
        out[:, xx] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear_u8(img, out_h, out_w):
    """ImagingResample for an 'L' image: horizontal pass (all rows), then vertical pass on its 8-bit result."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    tmp = _pil_pass(img, *pil_bilinear_coeffs(img.shape[1], out_w))
    return np.ascontiguousarray(_pil_pass(np.ascontiguousarray(tmp.T), *pil_bilinear_coeffs(img.shape[0], out_h)).T)


def _paste(small, bbox, image_shape):
    y1, x1, y2, x2 = (int(v) for v in bbox)
    full = np.zeros(tuple(image_shape[:2]), np.uint8)
    if y2 > y1 and x2 > x1:
        full[y1:y2, x1:x2] = small
    return full


def unmold_mask(mask, bbox, image_shape):
    """utils.py:447-465 on the restated resize: bytescale -> bilinear resize to the box -> v / 255 >= 0.5 (v >= 128)
    -> paste.  An empty box pastes nothing (the reference drops zero-area detections first, model.py:786-795)."""
    y1, x1, y2, x2 = (int(v) for v in bbox)
    m = np.squeeze(np.asarray(mask, dtype=np.float32))
    if y2 <= y1 or x2 <= x1:
        return np.zeros(tuple(image_shape[:2]), np.uint8)
    r = pil_resize_bilinear_u8(bytescale_f32(m), y2 - y1, x2 - x1)
    small = np.where(r.astype(np.float32) / np.float32(255.0) >= 0.5, 1, 0).astype(np.uint8)
    return _paste(small, bbox, image_shape)


def unmold_mask_pil(mask, bbox, image_shape):
    """The same function through the real Pillow (what scipy.misc.imresize calls): toimage -> Image.resize(BILINEAR)
    -> fromimage.  Used to pin unmold_mask(); needs PIL."""
    from PIL import Image
    y1, x1, y2, x2 = (int(v) for v in bbox)
    m = np.squeeze(np.asarray(mask, dtype=np.float32))
    if y2 <= y1 or x2 <= x1:
        return np.zeros(tuple(image_shape[:2]), np.uint8)
    b = bytescale_f32(m)
    im = Image.frombytes("L", (b.shape[1], b.shape[0]), b.tobytes())
    r = np.asarray(im.resize((x2 - x1, y2 - y1), resample=Image.BILINEAR), dtype=np.uint8)
    small = np.where(r.astype(np.float32) / np.float32(255.0) >= 0.5, 1, 0).astype(np.uint8)
    return _paste(small, bbox, image_shape)


# ---------------------------------------------------------------------------
# RPN output re-layout: RPN.forward's permute / view / softmax (modal/modals.py:
# 394-410) and the concatenation over levels (model.py:553-563), restated with numpy index arithmetic.
# ---------------------------------------------------------------------------
def rpn_pack(class_maps, bbox_maps):
    """class_maps[l] f32 [B,2a,H,W], bbox_maps[l] f32 [B,4a,H,W] -> (logits [B,A,2], probs [B,A,2], bbox [B,A,4])."""
    lg, bx = [], []
    for c, b in zip(class_maps, bbox_maps):
        c = np.asarray(c, np.float32)
        b = np.asarray(b, np.float32)
        B = c.shape[0]
        lg.append(np.ascontiguousarray(c.transpose(0, 2, 3, 1)).reshape(B, -1, 2))        # modals.py:394-397
        bx.append(np.ascontiguousarray(b.transpose(0, 2, 3, 1)).reshape(B, -1, 4))        # modals.py:407-410
    logits = np.concatenate(lg, axis=1)                                                   # model.py:560
    bbox = np.concatenate(bx, axis=1)
    m = logits.max(axis=2, keepdims=True)
    e = np.exp((logits - m).astype(np.float32)).astype(np.float32)                        # Softmax(dim=2), modals.py:400
    probs = (e / (e[..., :1] + e[..., 1:]).astype(np.float32)).astype(np.float32)
    return logits, probs, bbox


def resize_image(image, out_hw):
    """utils.resize_image's scipy.misc.imresize(image, (max_dim, max_dim)) (utils.py:352) for a uint8 image [h, w] or
    [h, w, C]: no bytescale for uint8 input, Pillow's 8-bit bilinear resample per band (pil_resize_bilinear_u8)."""
    a = np.asarray(image)
    assert a.dtype == np.uint8
    if a.ndim == 2:
        return pil_resize_bilinear_u8(a, int(out_hw[0]), int(out_hw[1]))
    return np.stack([pil_resize_bilinear_u8(np.ascontiguousarray(a[:, :, c]), int(out_hw[0]), int(out_hw[1]))
                     for c in range(a.shape[2])], axis=2)


def resize_image_pil(image, out_hw):
    """The same through the real Pillow (toimage of a uint8 [h, w, 3] array is mode 'RGB'); pins resize_image()."""
    from PIL import Image
    a = np.ascontiguousarray(image)
    mode = "L" if a.ndim == 2 else {3: "RGB", 4: "RGBA"}[a.shape[2]]
    im = Image.frombytes(mode, (a.shape[1], a.shape[0]), a.tobytes())
    return np.asarray(im.resize((int(out_hw[1]), int(out_hw[0])), resample=Image.BILINEAR), dtype=np.uint8)


# ---------------------------------------------------------------------------
# Speed-only comparator: the reference's CUDA kernels, unmodified, recompiled for sm_100a (oracle/_ref/libref_cuda.so;
# BASELINE.md section 4).  NOT a parity oracle (FMA contraction, `>` instead of `>=`, atomicAdd order).  The callers pass
# raw device pointers (torch tensors' data_ptr()); nothing here touches the product path.
# ---------------------------------------------------------------------------
_ref_cuda = None


def ref_cuda_available() -> bool:
    return os.path.exists(_REF_CUDA_SO)


def _ref_cuda_lib():
    global _ref_cuda
    if _ref_cuda is None:
        _ref_cuda = C.CDLL(_REF_CUDA_SO)
        _ref_cuda.CropAndResizeLaucher.restype = None
        _ref_cuda.CropAndResizeBackpropImageLaucher.restype = None
        _ref_cuda._nms.restype = None
    return _ref_cuda


def ref_cuda_crop_fwd(image_ptr, boxes_ptr, ind_ptr, n, B, H, W, ph, pw, depth, ext, crops_ptr, stream=0):
    """CropAndResizeLaucher (cuda/crop_and_resize_kernel.cu:166-192): NCHW image, crops [n, depth, ph, pw]."""
    _ref_cuda_lib().CropAndResizeLaucher(C.c_void_p(image_ptr), C.c_void_p(boxes_ptr), C.c_void_p(ind_ptr), C.c_int(n), C.c_int(B),
                                         C.c_int(H), C.c_int(W), C.c_int(ph), C.c_int(pw), C.c_int(depth), C.c_float(ext),
                                         C.c_void_p(crops_ptr), C.c_void_p(stream))


def ref_cuda_crop_bwd(grads_ptr, boxes_ptr, ind_ptr, n, B, H, W, ph, pw, depth, grads_image_ptr, stream=0):
    """CropAndResizeBackpropImageLaucher (crop_and_resize_kernel.cu:195-220): atomicAdd scatter into a ZEROED image."""
    _ref_cuda_lib().CropAndResizeBackpropImageLaucher(C.c_void_p(grads_ptr), C.c_void_p(boxes_ptr), C.c_void_p(ind_ptr), C.c_int(n),
                                                      C.c_int(B), C.c_int(H), C.c_int(W), C.c_int(ph), C.c_int(pw), C.c_int(depth),
                                                      C.c_void_p(grads_image_ptr), C.c_void_p(stream))


def ref_cuda_nms_mask(n, boxes_ptr, mask_ptr, thresh):
    """_nms (cuda/nms_kernel.cu:73-84): the n x ceil(n/64) IoU bit matrix of boxes sorted by score, default stream."""
    _ref_cuda_lib()._nms(C.c_int(n), C.c_void_p(boxes_ptr), C.c_void_p(mask_ptr), C.c_float(thresh))


def nms_mask_scan(mask):
    """Host scan of the bit matrix (restatement of nms_cuda.c:33-58): positions of the survivors in the sorted order."""
    mask = np.ascontiguousarray(mask, np.uint64)
    n = mask.shape[0]
    keep = np.empty(n, np.int64)
    k = _lib().orc_nms_mask_scan(mask.ctypes.data_as(C.c_void_p), C.c_int(n), keep.ctypes.data_as(C.c_void_p))
    return keep[:k]


# ---------------------------------------------------------------------------
# restatement: load_image_gt (modal/Functions.py:675-736) from the pieces above
# ---------------------------------------------------------------------------
def load_image_gt(label, image, num_classes, max_dim, augment, seed, image_id=0):
    """load_image_gt on an in-memory (label map, image) pair: AmodalDataset.load_layer2 (layer_decode), utils.resize_image
    (resize_image: Pillow-exact), utils.resize_layer (scipy.ndimage.zoom order 0 -- the library the reference calls), the
    seeded flip (`random.randint`), utils.extract_bboxes' jitter (`np.random.rand`, 4 per instance), compose_image_meta
    and the final [H,W,n,L] uint8 layout.  Seeds both generators like tests/golden/make_golden_loadgt.py does."""
    import random
    import scipy.ndimage
    random.seed(seed)
    np.random.seed(seed)
    L = num_classes - 1
    planes, n = layer_decode(label, L, n_max=32)                 # [32, L, H, W]
    mask_layers = planes[:n].transpose(2, 3, 1, 0).astype(bool)    # [H, W, L, n] as load_layer2 returns it
    shape = image.shape
    h, w = shape[:2]
    img = resize_image(image, (max_dim, max_dim))
    scale = (max_dim / h, max_dim / w)
    window = (0, 0, max_dim, max_dim)
    mask_layers = scipy.ndimage.zoom(mask_layers, zoom=[scale[0], scale[1], 1, 1], order=0)
    if augment and random.randint(0, 1):
        img = np.fliplr(img)
        mask_layers = np.fliplr(mask_layers)
    amodal = np.sum(mask_layers, axis=2)
    boxes = np.zeros([amodal.shape[-1], 4], dtype=np.int32)
    for i in range(amodal.shape[-1]):
        m = amodal[:, :, i]
        hz = np.where(np.any(m, axis=0))[0]
        vt = np.where(np.any(m, axis=1))[0]
        if hz.shape[0]:
            x1, x2 = hz[[0, -1]]
            y1, y2 = vt[[0, -1]]
            x2 += 1
            y2 += 1
        else:
            x1, x2, y1, y2 = 0, 0, 0, 0
        box = np.array([y1, x1, y2, x2]) + (np.random.rand(4) * 2 - 1) * (y2 - y1, x2 - x1, y2 - y1, x2 - x1) / 15
        box[box < 0] = 0
        boxes[i] = box
    meta = np.array([image_id] + list(shape) + list(window) + [1] * 128)
    return img, meta, np.ones(n, np.int32), boxes.astype(np.int32), (np.swapaxes(mask_layers, 2, 3) > 0).astype(np.uint8)

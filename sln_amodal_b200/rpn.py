"""RPN output re-layout in front of proposal_layer.

    rpn_pack(class_maps, bbox_maps)     -> [rpn_class_logits [B,A,2], rpn_probs [B,A,2], rpn_bbox [B,A,4]]
    rpn_forward_levels(rpn, feature_maps) -> the same list, driving the reference's RPN module level by level

Reference: RPN.forward (modal/modals.py:388-412) permutes, copies, views and soft-maxes each level's two conv outputs,
and MaskRCNN.predict (model.py:553-563) concatenates the five levels: 10 permute copies + 5 softmax + 3 cat launches per
step.  Here the conv outputs of all levels go through ONE launch (sln_rpn_pack); the backward of the logits / deltas
(rpn losses, model.py:423-436) is one launch too (sln_rpn_unpack_grads).  rpn_probs is not differentiable here -- the
reference only consumes it through proposal_layer, which works on .data (Functions.py:128-131).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _level_args(maps):
    nl = len(maps)
    hs = (C.c_int * nl)(*[int(m.shape[2]) for m in maps])
    ws = (C.c_int * nl)(*[int(m.shape[3]) for m in maps])
    return nl, hs, ws


def _ptrs(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _is_nhwc(t):
    return t.dim() == 4 and t.shape[1] > 1 and t.stride(1) == 1 and not t.is_contiguous()


def _uniform_layout(maps):
    """-> (tensors, nhwc flag): all levels dense in one layout (channels_last kept as it is, anything else -> NCHW)"""
    if all(m.is_contiguous() for m in maps):
        return list(maps), False
    if all(_is_nhwc(m) and m.is_contiguous(memory_format=torch.channels_last) for m in maps):
        return list(maps), True
    return [m.contiguous() for m in maps], False


def _pack_forward(a, want_probs, n_levels, maps):
    """-> (logits, probs or None, bbox, nhwc flag, dense tensors): the one launch"""
    B = maps[0].shape[0]
    for l in range(n_levels):
        c, b = maps[l], maps[n_levels + l]
        if not (c.is_cuda and b.is_cuda) or c.dtype != torch.float32 or b.dtype != torch.float32 or c.dim() != 4 or b.dim() != 4:
            raise _lib.SlnError("rpn_pack: float32 CUDA tensors [B, C, H, W] expected (there is no CPU fallback)")
        if c.shape[1] != 2 * a or b.shape[1] != 4 * a or c.shape[2:] != b.shape[2:] or c.shape[0] != B or b.shape[0] != B:
            raise _lib.SlnError("rpn_pack: level shapes must be [B, 2a, H, W] / [B, 4a, H, W]")
    tens, nhwc = _uniform_layout(maps)
    nl, hs, ws = _level_args(tens[:n_levels])
    A = a * sum(int(m.shape[2]) * int(m.shape[3]) for m in tens[:n_levels])
    dev = tens[0].device
    logits = torch.empty((B, A, 2), dtype=torch.float32, device=dev)
    probs = torch.empty((B, A, 2), dtype=torch.float32, device=dev) if want_probs else None
    bbox = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
    if B and A:
        with torch.cuda.device(dev):
            check(lib().sln_rpn_pack(_ptrs(tens[:n_levels]), _ptrs(tens[n_levels:]), hs, ws, nl, B, a,
                                     _lib.LAYOUT_NHWC if nhwc else _lib.LAYOUT_NCHW, ptr(logits), ptr(probs), ptr(bbox), stream_ptr()),
                  "sln_rpn_pack")
        _lib.count_launches(1)
    return logits, probs, bbox, nhwc, tens


class _RpnPack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, want_probs, n_levels, *maps):
        logits, probs, bbox, nhwc, tens = _pack_forward(a, want_probs, n_levels, list(maps))
        ctx.a, ctx.n_levels, ctx.nhwc = a, n_levels, nhwc
        ctx.shapes = [tuple(m.shape) for m in tens]
        if probs is None:
            probs = logits.new_empty(0)
        ctx.mark_non_differentiable(probs)
        return logits, probs, bbox

    @staticmethod
    def backward(ctx, g_logits, g_probs, g_bbox):
        n = ctx.n_levels
        dev = (g_logits if g_logits is not None else g_bbox).device
        fmt = torch.channels_last if ctx.nhwc else torch.contiguous_format
        grads = [torch.empty(s, dtype=torch.float32, device=dev, memory_format=fmt) for s in ctx.shapes]
        gl = g_logits.contiguous() if g_logits is not None else None
        gb = g_bbox.contiguous() if g_bbox is not None else None
        B = ctx.shapes[0][0]
        if B and sum(s[2] * s[3] for s in ctx.shapes[:n]):
            hs = (C.c_int * n)(*[s[2] for s in ctx.shapes[:n]])
            ws = (C.c_int * n)(*[s[3] for s in ctx.shapes[:n]])
            with torch.cuda.device(dev):
                check(lib().sln_rpn_unpack_grads(ptr(gl), ptr(gb), hs, ws, n, B, ctx.a, _lib.LAYOUT_NHWC if ctx.nhwc else _lib.LAYOUT_NCHW,
                                                 _ptrs(grads[:n]), _ptrs(grads[n:]), stream_ptr()), "sln_rpn_unpack_grads")
            _lib.count_launches(1)
        return (None, None, None) + tuple(grads)


def rpn_pack(class_maps, bbox_maps, anchors_per_location=None, want_probs=True):
    """class_maps[l] [B, 2a, H_l, W_l], bbox_maps[l] [B, 4a, H_l, W_l] (conv outputs, NCHW or channels_last) ->
    [rpn_class_logits [B,A,2], rpn_probs [B,A,2], rpn_bbox [B,A,4]] concatenated over the levels in list order."""
    if len(class_maps) != len(bbox_maps) or not class_maps:
        raise _lib.SlnError("rpn_pack: one class map and one bbox map per level")
    a = int(anchors_per_location) if anchors_per_location else int(class_maps[0].shape[1]) // 2
    maps = list(class_maps) + list(bbox_maps)
    if not (torch.is_grad_enabled() and any(m.requires_grad for m in maps)):        # inference: no autograd node
        logits, probs, bbox = _pack_forward(a, bool(want_probs), len(class_maps), maps)[:3]
        return [logits, probs, bbox]
    logits, probs, bbox = _RpnPack.apply(a, bool(want_probs), len(class_maps), *maps)
    return [logits, probs if want_probs else None, bbox]


def rpn_forward_levels(rpn, feature_maps):
    """The loop + concatenation of MaskRCNN.predict (model.py:553-563) on the reference's RPN module (modal/modals.py:
    369-412): its own convolutions per level (stock cuDNN, out of scope), then one re-layout launch for everything."""
    cls_maps, box_maps = [], []
    for p in feature_maps:
        x = rpn.relu(rpn.conv_shared(rpn.padding(p)))
        cls_maps.append(rpn.conv_class(x))
        box_maps.append(rpn.conv_bbox(x))
    return rpn_pack(cls_maps, box_maps)

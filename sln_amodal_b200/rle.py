"""COCO run-length encoding of binary masks on the device -- the codec after the path.

    encode(masks) -> [{"size": [h, w], "counts": bytes}, ...]        (pycocotools.mask.encode's result)

The reference's evaluation (amodal_train.py:371-400) encodes each full-resolution detection mask on the host with
cocoapi's rleEncode + rleToString (cocoapi/common/maskApi.c:32-41, 204-216).  Here the run boundaries of all masks are
found in one launch (ops.rle_counts_device); only the run lengths (a few KB per mask) travel to the host, where the
compressed string is produced by the C helper sln_rle_to_string.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def rle_counts_device(planes, cap=None):
    """planes u8 / bool [n, a] (CUDA), each row in the memory order to encode -> (counts u32-as-int32 [n, cap], m int32 [n]).
    m[i] < 0 means mask i has -m[i] runs > cap."""
    if not planes.is_cuda:
        raise _lib.SlnError("planes must be a CUDA tensor (there is no CPU fallback)")
    if planes.dtype == torch.bool:
        planes = planes.view(torch.uint8)
    if planes.dtype != torch.uint8 or planes.dim() != 2:
        raise _lib.SlnError("planes must be uint8 / bool [n, a]")
    planes = planes.contiguous()
    n, a = planes.shape
    if cap is None:
        cap = int(min(a + 1, 1 << 16))
    counts = torch.empty((n, cap), dtype=torch.int32, device=planes.device)
    m = torch.empty(n, dtype=torch.int32, device=planes.device)
    if n:
        with torch.cuda.device(planes.device):
            check(lib().sln_rle_encode(ptr(planes), n, a, ptr(counts), cap, ptr(m), stream_ptr()), "sln_rle_encode")
        _lib.count_launches(1)
    return counts, m


def counts_to_string(counts):
    """maskApi.c:204-216 for one list of run lengths (numpy uint32) -> bytes."""
    c = np.ascontiguousarray(counts, dtype=np.uint32)
    cap = 6 * int(c.size) + 1
    buf = C.create_string_buffer(cap)
    n = lib().sln_rle_to_string(c.ctypes.data_as(C.c_void_p), int(c.size), buf, cap)
    if n < 0:
        raise _lib.SlnError("rle string buffer too small")
    return buf.raw[:n]


def encode(masks):
    """masks: CUDA tensor [n, h, w] (u8 / bool planes) or numpy / tensor [h, w, n] like pycocotools.mask.encode takes.
    Returns one {"size": [h, w], "counts": bytes} per mask; runs are taken in column-major order like pycocotools."""
    if isinstance(masks, np.ndarray):
        masks = torch.from_numpy(np.ascontiguousarray(np.moveaxis(masks, -1, 0))).cuda()
    if masks.dim() != 3:
        raise _lib.SlnError("masks must be [n, h, w] (device planes) or a numpy [h, w, n]")
    n, h, w = masks.shape
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    cols = masks.transpose(1, 2).contiguous().view(n, h * w)           # column-major planes
    counts, m = rle_counts_device(cols)
    m_h = m.cpu().numpy()
    if (m_h < 0).any():                                                # rare: more runs than the default capacity
        counts, m = rle_counts_device(cols, cap=int(-m_h.min()))
        m_h = m.cpu().numpy()
    kmax = int(m_h.max()) if n else 0
    c_h = counts[:, :kmax].cpu().numpy().view(np.uint32)
    return [{"size": [int(h), int(w)], "counts": counts_to_string(c_h[i, : m_h[i]])} for i in range(n)]

"""`.npz['layer']` reader (the format on the input side of the path).

The reference reads every image's bit-packed instance label map with `np.load(path)['layer']`
(amodal_train.py:238): a zip archive whose member `layer.npy` is stored or deflated, a `.npy` header, then H*W
little-endian uint64 words.  numpy inflates into a fresh pageable array that the caller then copies again on its way to
the device.  Here the member is inflated straight into ONE pinned host buffer (no intermediate array) and handed to the
device with an asynchronous copy on the current stream, so the next image's inflate overlaps it; the zip / deflate /
npy-header handling is the standard library's (`zipfile`, `zlib`) and numpy's own header parser -- a host codec, as the
survey says, not a GPU workload.  Output is bit-identical to `np.load(path)[name]` (tests/test_cpu_host.py).
"""
from __future__ import annotations

import zipfile

import numpy as np
import torch


def read_member(path, name="layer", pinned=None):
    """Array `name` of the .npz at `path` as a CPU tensor backed by pinned memory (pinned=None: when CUDA is there).
    uint64 data comes back as int64 bit patterns (what the layer decoder takes); other dtypes as themselves."""
    if pinned is None:
        pinned = torch.cuda.is_available()
    with zipfile.ZipFile(path) as z:
        member = name if name in z.namelist() else name + ".npy"
        with z.open(member) as f:
            major, minor = np.lib.format.read_magic(f)
            if (major, minor) == (1, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_1_0(f)
            elif (major, minor) == (2, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_2_0(f)
            else:
                raise ValueError("unsupported .npy version %d.%d in %s" % (major, minor, path))
            if dtype.hasobject:
                raise ValueError("object arrays are not read (np.load would need allow_pickle)")
            if dtype.byteorder == ">":
                raise ValueError("big-endian .npy members are not supported")
            tdtype = {np.dtype("uint64"): torch.int64}.get(dtype) or torch.from_numpy(np.empty(0, dtype)).dtype
            n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
            store_shape = tuple(reversed(shape)) if fortran else tuple(shape)
            buf = torch.empty(store_shape, dtype=tdtype, pin_memory=bool(pinned))
            view = memoryview(buf.numpy()).cast("B") if n else memoryview(b"")
            got = 0
            while got < len(view):                     # ZipExtFile.readinto inflates into the buffer it is given
                k = f.readinto(view[got:])
                if not k:
                    raise ValueError("truncated member %s in %s" % (member, path))
                got += k
    if fortran:
        buf = buf.permute(*reversed(range(buf.dim())))
    return buf


def load_layer_label(path, device=None, name="layer"):
    """The image's label map as an int64 [H, W] tensor on `device` (default: the current CUDA device): inflate into
    pinned memory, one asynchronous copy.  Same bits as torch.from_numpy(np.load(path)['layer'].view(np.int64))."""
    host = read_member(path, name)
    if host.dtype != torch.int64:
        raise TypeError("layer label maps are 64-bit integers (uint64 on disk, modal/Functions.py:1012-1095), got %s" % host.dtype)
    device = torch.device("cuda") if device is None else torch.device(device)
    if device.type != "cuda":
        return host
    return host.contiguous().to(device, non_blocking=True)

"""Semantics-aware distance-map target encoding on the device.

* decode_layers / load_layer2: drop-in for AmodalDataset.load_layer2 (reference
  amodal_train.py:236-271) and the codec it drives (modal/Functions.py:1012-1095).
* sem_dist_targets: the batched device path -- layer planes + exact squared EDT of every
  plane (the EDT is an addition of BASELINE.json's north star; the reference has none).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _as_label_tensor(label, device):
    if isinstance(label, np.ndarray):
        if label.dtype.itemsize != 8 or label.dtype.kind not in "ui":
            raise TypeError("layer label maps are 64-bit integers (uint64 on disk, Functions.py:1012-1095), got %s" % label.dtype)
        label = torch.from_numpy(np.ascontiguousarray(label).view(np.int64))
    elif label.dtype != torch.int64:
        raise TypeError("layer label tensors must be int64 (the bits of the uint64 map), got %s" % label.dtype)
    return label.to(device)


def decode_layers(label, num_classes, n_max=32):
    """label: u64 [H,W] or [B,H,W] (numpy or tensor) -> (planes u8 [B,n_max,L,H,W], n_obj int32[B])
    on the device, L = num_classes - 1 (amodal_train.py:246)."""
    dev = label.device if isinstance(label, torch.Tensor) and label.is_cuda else torch.device("cuda")
    t = _as_label_tensor(label, dev)
    return ops.layer_decode_device(t, num_classes - 1, n_max)


def load_layer2(dataset, image_id, config):
    """Same contract as AmodalDataset.load_layer2 (amodal_train.py:236-271): reads the image's
    `<path>.npz['layer']`, returns (mask_layers bool [H,W,L,n_obj], class_ids int32 [n_obj]) as
    numpy arrays; falls through to dataset.load_mask's empty result when no object decodes."""
    from . import npz
    image_info = dataset.image_info[image_id]
    layer = npz.load_layer_label(image_info['path'][:-4] + '.npz')       # inflate into pinned memory, asynchronous copy
    planes, n_obj = decode_layers(layer, config.NUM_CLASSES, n_max=32)
    n = int(n_obj[0].item())
    if n == 0:
        # the reference calls super(AmodalDataset, self).load_mask (amodal_train.py:268-270): the empty mask of utils.Dataset.
        # install() binds this function onto AmodalDataset, so "the class that owns load_layer2" is found in the MRO --
        # for an instance of a SUBCLASS, super(type(dataset), ...) would land on AmodalDataset.load_mask instead.
        owner = next((k for k in type(dataset).__mro__ if "load_layer2" in k.__dict__), type(dataset))
        return super(owner, dataset).load_mask(image_id)
    mask_layers = planes[0, :n].permute(2, 3, 1, 0).contiguous().cpu().numpy().astype(bool)
    return mask_layers, np.ones(n, dtype=np.int32)


def sem_dist_targets(label, num_classes, n_max=20):
    """label [B,H,W] -> dict(layers u8 [B,n_max,L,H,W], n_obj int32[B], dist_sq i32 [B,n_max,L,H,W]).
    dist_sq = squared Euclidean distance of every pixel to the nearest zero pixel of its plane."""
    planes, n_obj = decode_layers(label, num_classes, n_max=n_max)
    dist = ops.edt_sq_device(planes)
    return {"layers": planes, "n_obj": n_obj, "dist_sq": dist}

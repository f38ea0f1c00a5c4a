"""install(): route the reference's unmodified model code through this library.

Import-path shadowing: this repo ships top-level packages `roialign.roi_align.crop_and_resize`
and `nms.nms_wrapper` (same import paths as the reference, modal/modals.py:6,
modal/Functions.py:5,7).  With this repo ahead of the reference on sys.path, the reference's
`from roialign.roi_align.crop_and_resize import CropAndResizeFunction` and
`from nms.nms_wrapper import nms` resolve here without touching a reference file.

install() additionally rebinds the Python-level functions of the path to the fused versions.
The reference looks them up in module globals at call time (model.py:570,583,593,646 via
`from modal.Functions import *`; Classifier/Mask look pyramid_roi_align up in modal.modals).
"""
from __future__ import annotations

import sys


def install(model_module=None, functions_module=None, modals_module=None, dataset_class=None,
            channels_last_model=None):
    """Rebind the hot-path functions in the already-imported reference modules.

    model_module / functions_module / modals_module: the reference's `model`, `modal.Functions`,
    `modal.modals` modules (default: whatever is in sys.modules).  dataset_class: the reference's
    AmodalDataset (gets the device layer decoder as load_layer2).  channels_last_model: an
    nn.Module to convert to channels_last so FPN outputs feed the NHWC kernels natively.
    Returns the list of rebinding performed (for logging / tests)."""
    from . import proposal, pyramid, semdist, detection, targets

    model_module = model_module or sys.modules.get("model")
    functions_module = functions_module or sys.modules.get("modal.Functions")
    modals_module = modals_module or sys.modules.get("modal.modals")
    done = []

    def bind(mod, name, fn):
        if mod is not None and hasattr(mod, name):
            setattr(mod, name, fn)
            done.append(f"{mod.__name__}.{name}")

    for mod in (model_module, functions_module):
        bind(mod, "proposal_layer", proposal.proposal_layer)
        bind(mod, "refine_detections", detection.refine_detections)
        bind(mod, "detection_target_layer", targets.detection_target_layer)
        bind(mod, "bbox_overlaps", targets.bbox_overlaps)
        bind(mod, "build_rpn_targets", targets.build_rpn_targets)
        bind(mod, "load_image_gt", targets.load_image_gt)                     # model.py:80 looks it up in its own globals
        bind(mod, "pyramid_roi_align_image", pyramid.pyramid_roi_align_image)
    for mod in (modals_module, model_module):
        bind(mod, "pyramid_roi_align", pyramid.pyramid_roi_align)
        bind(mod, "pyramid_roi_align_image", pyramid.pyramid_roi_align_image)
    if model_module is not None and hasattr(getattr(model_module, "MaskRCNN", None), "unmold_detections"):
        from . import unmold

        def _unmold_detections(self, detections, mrcnn_mask, image_shape, window):
            return unmold.unmold_detections(detections, mrcnn_mask, image_shape, window)

        model_module.MaskRCNN.unmold_detections = _unmold_detections       # model.py:747-806
        done.append(f"{model_module.__name__}.MaskRCNN.unmold_detections")
    if dataset_class is not None:
        dataset_class.load_layer2 = semdist.load_layer2
        done.append(f"{dataset_class.__name__}.load_layer2")
    if channels_last_model is not None:
        import torch
        channels_last_model.to(memory_format=torch.channels_last)
        done.append("model.to(channels_last)")
    return done

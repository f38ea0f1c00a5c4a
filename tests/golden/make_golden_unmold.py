"""Generates tests/golden/unmold_detections.npz and resize_image.npz by running the REFERENCE's own Python, unmodified,
imported from /root/reference: MaskRCNN.unmold_detections (model.py:747-806) with utils.unmold_mask (utils.py:447-465),
and utils.resize_image (utils.py:301-356).  Both call scipy.misc.imresize, which is gone from the scipy installed here
(removed in 1.3); the generator provides it as a module-level shim that restates scipy 1.0's pilutil (bytescale ->
toimage -> PIL.Image.resize -> fromimage) around the REAL Pillow.  Everything else -- the box arithmetic, the zero-area
filter, the threshold, the paste, the stacking, resize_image's window / scale / padding -- is the reference's own code.
Run in the build container only:

    python tests/golden/make_golden_unmold.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import setup_reference_imports  # noqa: E402


def _imresize(arr, size, interp="bilinear", mode=None):
    """scipy 1.0 misc/pilutil.py imresize for the two cases the reference uses: float 2-D masks (bytescale first) and
    uint8 images; `size` is (rows, cols) as in the reference's calls."""
    from PIL import Image
    from oracle import oracle
    a = np.asarray(arr)
    func = {"nearest": 0, "lanczos": 1, "bilinear": 2, "bicubic": 3, "cubic": 3}[interp]
    if a.ndim == 2:
        b = a if a.dtype == np.uint8 else oracle.bytescale_f32(a.astype(np.float32))
        im = Image.frombytes("L", (b.shape[1], b.shape[0]), np.ascontiguousarray(b).tobytes())
    else:
        assert a.dtype == np.uint8 and a.shape[2] == 3
        im = Image.frombytes("RGB", (a.shape[1], a.shape[0]), np.ascontiguousarray(a).tobytes())
    return np.asarray(im.resize((int(size[1]), int(size[0])), resample=func))


def main():
    setup_reference_imports()
    import scipy
    if not hasattr(scipy, "misc") or not hasattr(getattr(scipy, "misc", None), "imresize"):
        misc = types.ModuleType("scipy.misc")
        misc.imresize = _imresize
        sys.modules["scipy.misc"] = misc
        scipy.misc = misc
    import utils as U                       # the reference's utils.py
    import model as M                       # the reference's model.py

    rng = np.random.default_rng(23)
    H, W = 96, 128
    N, pad = 7, 3
    det = np.zeros((N + pad, 6), np.float32)
    y1x1 = rng.uniform(0, 40, (N, 2))
    det[:N, 0:2] = y1x1
    det[:N, 2:4] = y1x1 + rng.uniform(1.5, 50, (N, 2))
    det[:N, 4] = rng.integers(1, 4, N)
    det[:N, 5] = rng.random(N)
    det[2, 2] = det[2, 0]                             # zero height: filtered out by the zero-area rule (model.py:786-795)
    mrcnn_mask = rng.random((N + pad, 28, 28, 2)).astype(np.float32)
    window = np.array([0, 0, 64, 96])                 # image part of a padded 64 x 96 network input
    image_shape = (H, W, 3)
    boxes, class_ids, scores, masks = M.MaskRCNN.unmold_detections(None, det, mrcnn_mask, image_shape, window)
    np.savez_compressed(os.path.join(HERE, "unmold_detections.npz"), detections=det, mrcnn_mask=mrcnn_mask,
                        window=window, image_shape=np.array(image_shape), boxes=boxes, class_ids=class_ids,
                        scores=scores, masks=masks.astype(np.uint8))
    print("unmold_detections:", boxes.shape, masks.shape, masks.dtype, int(masks.sum()))

    img = rng.integers(0, 256, (75, 50, 3)).astype(np.uint8)
    out, win, scale, padding = U.resize_image(img, min_dim=None, max_dim=64, padding=False)
    np.savez_compressed(os.path.join(HERE, "resize_image.npz"), image=img, resized=out, window=np.array(win),
                        scale=np.array(scale), padding=np.array(padding))
    print("resize_image:", out.shape, win, scale)


if __name__ == "__main__":
    main()

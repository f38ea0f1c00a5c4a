"""Generates tests/golden/load_image_gt.npz by running the REFERENCE's own load_image_gt (modal/Functions.py:675-736),
unmodified, imported from /root/reference, on a two-image synthetic dataset: AmodalDataset.load_layer2 reading a real
`.npz['layer']` file (amodal_train.py:236-271), utils.resize_image, utils.resize_layer (scipy.ndimage.zoom), the random
flip (`random.randint`), utils.extract_bboxes with its `np.random.rand` jitter, compose_image_meta and the final
swapaxes / uint8 cast.  scipy.misc.imresize is supplied around the real Pillow exactly as in make_golden_unmold.py.
The seeds of both generators are stored with the outputs: a replay that draws in the same order reproduces them.
Run in the build container only:

    python tests/golden/make_golden_loadgt.py
"""
import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import setup_reference_imports  # noqa: E402
from make_golden_unmold import _imresize  # noqa: E402


class Cfg:
    IMAGE_MIN_DIM = 48
    IMAGE_MAX_DIM = 64
    IMAGE_PADDING = True
    USE_MINI_MASK = False
    MINI_MASK_SHAPE = (56, 56)
    NUM_CLASSES = 3


def main():
    setup_reference_imports()
    import scipy
    if not hasattr(scipy, "misc") or not hasattr(getattr(scipy, "misc", None), "imresize"):
        misc = types.ModuleType("scipy.misc")
        misc.imresize = _imresize
        sys.modules["scipy.misc"] = misc
        scipy.misc = misc
    import amodal_train as AT
    from modal import Functions as F
    from sln_amodal_b200 import synth

    tmp = tempfile.mkdtemp(prefix="sln_loadgt_")
    rng = np.random.default_rng(77)
    out = {}
    cases = [(75, 100, 5, 3, False, 11), (96, 64, 7, 4, True, 12), (50, 80, 4, 2, True, 13)]
    for i, (H, W, n_obj, num_classes, augment, seed) in enumerate(cases):
        label = synth.label_map(H, W, n=n_obj, seed=400 + i, min_piece=10)
        img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        path = os.path.join(tmp, "img%d.jpg" % i)
        np.savez_compressed(path[:-4] + ".npz", layer=label)
        ds = AT.AmodalDataset.__new__(AT.AmodalDataset)
        ds.image_info = [{"path": path, "height": H, "width": W, "id": i}]
        ds.load_image = lambda image_id, _img=img: _img
        cfg = Cfg()
        cfg.NUM_CLASSES = num_classes
        random.seed(seed)
        np.random.seed(seed)
        image, meta, class_ids, bbox, masks = F.load_image_gt(ds, cfg, 0, augment=augment, use_mini_mask=False)
        out.update({"label%d" % i: label, "image_in%d" % i: img, "num_classes%d" % i: num_classes,
                    "augment%d" % i: int(augment), "seed%d" % i: seed, "image%d" % i: image, "meta%d" % i: meta,
                    "class_ids%d" % i: class_ids, "bbox%d" % i: bbox, "masks%d" % i: np.packbits(masks),
                    "masks_shape%d" % i: np.array(masks.shape)})
        print(i, image.shape, meta[:8], class_ids, bbox.tolist(), masks.shape, masks.dtype, int(masks.sum()))
    np.savez_compressed(os.path.join(HERE, "load_image_gt.npz"), n_cases=len(cases), max_dim=Cfg.IMAGE_MAX_DIM,
                        min_dim=Cfg.IMAGE_MIN_DIM, **out)


if __name__ == "__main__":
    main()

"""Generates tests/golden/*.npz by running the REFERENCE's own Python (unmodified, imported from
/root/reference) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference's Python functions on the hot path (modal/Functions.py::proposal_layer,
refine_detections; modal/modals.py::pyramid_roi_align, pyramid_roi_align_image;
amodal_train.py::AmodalDataset.load_layer2) are imported as they are.  Their two native
dependencies are bound to the reference's own C sources compiled into oracle/_ref (shim packages
`roialign.roi_align.crop_and_resize` and `nms.nms_wrapper` created in a temp dir), and the
third-party imports that are missing here (matplotlib, skimage, tensorboardX, pycocotools, tqdm)
are stubbed in sys.modules.  Nothing of the reference is copied: only inputs and the outputs it
produced are stored.  /root/reference does not exist on the GPU box, so tests read the fixtures.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def setup_reference_imports():
    sys.path.insert(0, ROOT)
    from oracle import oracle
    oracle.build()
    assert oracle.ref_available(), "oracle/_ref must be built from /root/reference first"

    shim = tempfile.mkdtemp(prefix="sln_ref_shim_")
    os.makedirs(os.path.join(shim, "roialign", "roi_align"))
    os.makedirs(os.path.join(shim, "nms"))
    for d in ("roialign", "roialign/roi_align", "nms"):
        open(os.path.join(shim, d, "__init__.py"), "w").close()
    with open(os.path.join(shim, "roialign", "roi_align", "crop_and_resize.py"), "w") as f:
        f.write(
            "import torch\nfrom oracle import oracle\n"
            "class CropAndResizeFunction(object):\n"
            "    def __init__(self, ch, cw, ext=0):\n        self.ch, self.cw, self.ext = ch, cw, ext\n"
            "    def __call__(self, image, boxes, box_ind):\n"
            "        out = oracle.ref_crop_and_resize_fwd(image.detach().numpy(), boxes.detach().numpy(),\n"
            "                                             box_ind.detach().numpy(), self.ch, self.cw, self.ext)\n"
            "        return torch.from_numpy(out)\n")
    with open(os.path.join(shim, "nms", "nms_wrapper.py"), "w") as f:
        f.write(
            "import torch\nfrom oracle import oracle\n"
            "def nms(dets, thresh):\n"
            "    # pth_nms.py:10-24 (CPU branch) on the reference's cpu_nms; scores in the fixtures are tie-free\n"
            "    return torch.from_numpy(oracle.ref_nms(dets.detach().numpy(), thresh))\n")
    # order matters: shims first, then this repo (for `oracle`), then the reference
    sys.path[:0] = [shim]
    sys.path.append(REF)
    # drop this repo's own shadow packages so the shims win
    for name in list(sys.modules):
        if name.split(".")[0] in ("roialign", "nms"):
            del sys.modules[name]

    cm = _stub("matplotlib.cm")
    _stub("matplotlib", cm=cm, use=lambda *a, **k: None)
    _stub("matplotlib.pyplot")
    _stub("matplotlib.patches")
    _stub("matplotlib.lines")
    _stub("skimage")
    for sub in ("color", "io", "morphology", "transform"):
        _stub("skimage." + sub)
    _stub("skimage.measure", label=None, regionprops=None)
    _stub("tensorboardX", SummaryWriter=lambda *a, **k: None)
    _stub("pycocotools")
    _stub("pycocotools.coco", COCO=object)
    _stub("pycocotools.cocoeval", COCOeval=object)
    _stub("pycocotools.mask")
    _stub("tqdm", tqdm=lambda x, *a, **k: x)
    return shim


class Cfg:
    RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    IMAGE_SHAPE = np.array([1024, 1024, 3])
    GPU_COUNT = 0
    USE_NMS = True
    DETECTION_MIN_CONFIDENCE = 0.3
    DETECTION_NMS_THRESHOLD = 0.3
    DETECTION_MAX_INSTANCES = 100
    NUM_CLASSES = 3


def smooth_map(C, H, W, phase):
    """Deterministic feature map without an RNG (so tests can rebuild it bit for bit)."""
    c = np.arange(C, dtype=np.float32)[:, None, None]
    y = np.arange(H, dtype=np.float32)[None, :, None]
    x = np.arange(W, dtype=np.float32)[None, None, :]
    v = np.sin(np.float32(0.37) * y + np.float32(0.11) * c + np.float32(phase)) * np.cos(np.float32(0.23) * x - np.float32(0.05) * c)
    return (v + np.float32(0.01) * (y * x % np.float32(7.0))).astype(np.float32)[None]


def main():
    setup_reference_imports()
    from sln_amodal_b200 import synth
    import modal.Functions as F
    import modal.modals as M
    out = {}

    # ---- proposal_layer (Functions.py:114-178): 8000 anchors > the 6000 pre-NMS limit
    rng = np.random.default_rng(20240)
    A = 8000
    anchors = synth.nms_boxes(A, seed=5, kind="rpn").astype(np.float32)
    fg = rng.permutation(np.linspace(0.0, 1.0, A)).astype(np.float32)            # tie-free
    probs = np.stack([1 - fg, fg], 1).astype(np.float32)
    deltas = (rng.standard_normal((A, 4)) * np.array([0.8, 0.8, 1.2, 1.2])).astype(np.float32)
    res = F.proposal_layer([torch.from_numpy(probs).unsqueeze(0), torch.from_numpy(deltas).unsqueeze(0)],
                           proposal_count=1000, nms_threshold=0.7, anchors=torch.from_numpy(anchors), config=Cfg())
    np.savez_compressed(os.path.join(HERE, "proposal_layer.npz"), probs=probs, deltas=deltas, anchors=anchors,
                        proposal_count=1000, nms_threshold=np.float32(0.7), out=res.numpy())
    out["proposal_layer"] = tuple(res.shape)

    # ---- pyramid_roi_align (modals.py:20-110) and pyramid_roi_align_image (:112-157)
    C = 4
    sides = (64, 32, 16, 8)                      # "P2..P5" of a 256^2 image
    maps = [smooth_map(C, s, s, 0.3 * i) for i, s in enumerate(sides)]
    boxes = synth.roi_boxes(96, seed=77)
    b64 = boxes.astype(np.float64)
    raw = 4 + np.log2(np.sqrt((b64[:, 2] - b64[:, 0]) * (b64[:, 3] - b64[:, 1])) / (224.0 / 256.0))
    boxes = boxes[np.abs(raw - np.floor(raw) - 0.5) > 1e-3]          # stay off exact level boundaries
    pooled = {}
    for pool in (7, 16):
        r = M.pyramid_roi_align([torch.from_numpy(boxes).unsqueeze(0)] + [torch.from_numpy(m) for m in maps], pool, (256, 256, 3))
        pooled[pool] = r.numpy()
    img_crop = M.pyramid_roi_align_image([torch.from_numpy(boxes).unsqueeze(0), torch.from_numpy(maps[1])], 8, (256, 256, 3))
    np.savez_compressed(os.path.join(HERE, "pyramid_roi_align.npz"), boxes=boxes, sides=np.array(sides), channels=C,
                        pooled7=pooled[7], pooled16=pooled[16], image_crop8=img_crop.numpy())
    out["pyramid_roi_align"] = tuple(pooled[7].shape)

    # ---- refine_detections with the per-class NMS loop (Functions.py:453-557)
    rng = np.random.default_rng(99)
    N, K = 400, 6
    rois = synth.roi_boxes(N, seed=123)
    logits = rng.standard_normal((N, K)).astype(np.float32) * 2
    p = np.exp(logits - logits.max(1, keepdims=True))
    p = (p / p.sum(1, keepdims=True)).astype(np.float32)
    # tie-free winning scores
    win = p.argmax(1)
    p[np.arange(N), win] += (rng.permutation(N).astype(np.float32) * np.float32(1e-6))
    d = (rng.standard_normal((N, K, 4)) * 0.5).astype(np.float32)
    window = np.array([0, 0, 1024, 1024], np.float32)
    det, keep = F.refine_detections(torch.from_numpy(rois), torch.from_numpy(p), torch.from_numpy(d), window, Cfg())
    np.savez_compressed(os.path.join(HERE, "refine_detections.npz"), rois=rois, probs=p, deltas=d, window=window,
                        detections=det.numpy(), keep=keep.numpy())
    out["refine_detections"] = tuple(det.shape)

    # ---- refine_detections on the reference's DEFAULT branch, USE_NMS = False (config.py:78; Functions.py:526-546):
    # top-100 by score.  Case a: more than 100 non-background ROIs (truncation); case b: fewer (no truncation).
    cfg = Cfg()
    cfg.USE_NMS = False
    nn = {}
    for tag, n_use in (("a", N), ("b", 90)):
        det, keep = F.refine_detections(torch.from_numpy(rois[:n_use]), torch.from_numpy(p[:n_use]),
                                        torch.from_numpy(d[:n_use]), window, cfg)
        nn["n_" + tag] = n_use
        nn["detections_" + tag] = det.numpy()
        nn["keep_" + tag] = keep.numpy()
    np.savez_compressed(os.path.join(HERE, "refine_detections_nonms.npz"), rois=rois, probs=p, deltas=d, window=window, **nn)
    out["refine_detections_nonms"] = (tuple(nn["detections_a"].shape), tuple(nn["detections_b"].shape))

    # ---- layer codec through AmodalDataset.load_layer2 (amodal_train.py:236-271)
    import amodal_train as AT
    tmp = tempfile.mkdtemp(prefix="sln_layers_")
    cases = {}
    for i, (n_obj, num_classes) in enumerate([(6, 2), (9, 3), (12, 4), (5, 6)]):
        label = synth.label_map(96, 128, n=n_obj, seed=300 + i, min_piece=12)
        if i == 1:                                    # overlapping annotations / occluded-only piece
            label[0, :6] = (1 << 0) | (1 << 1)
            label[1, :6] = np.uint64(1 << 34)
        path = os.path.join(tmp, "img%d.jpg" % i)
        np.savez(path[:-4] + ".npz", layer=label)
        ds = AT.AmodalDataset.__new__(AT.AmodalDataset)
        ds.image_info = [{"path": path, "height": label.shape[0], "width": label.shape[1]}]
        cfg = Cfg()
        cfg.NUM_CLASSES = num_classes
        mask_layers, class_ids = ds.load_layer2(0, cfg)
        cases["label%d" % i] = label
        cases["num_classes%d" % i] = num_classes
        cases["mask_layers%d" % i] = np.packbits(mask_layers.astype(np.uint8))
        cases["shape%d" % i] = np.array(mask_layers.shape)
        cases["class_ids%d" % i] = class_ids
    np.savez_compressed(os.path.join(HERE, "load_layer2.npz"), n_cases=4, **cases)
    out["load_layer2"] = [tuple(cases["shape%d" % i]) for i in range(4)]
    print("golden fixtures written:", out)


if __name__ == "__main__":
    main()

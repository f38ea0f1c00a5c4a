"""Golden fixtures produced by the reference's own Python (tests/golden/make_golden.py, run in the
build container where /root/reference exists).  CPU tests pin the oracle restatement to them;
GPU tests hold the CUDA path to the same vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def smooth_map(C, H, W, phase):
    c = np.arange(C, dtype=np.float32)[:, None, None]
    y = np.arange(H, dtype=np.float32)[None, :, None]
    x = np.arange(W, dtype=np.float32)[None, None, :]
    v = np.sin(np.float32(0.37) * y + np.float32(0.11) * c + np.float32(phase)) * np.cos(np.float32(0.23) * x - np.float32(0.05) * c)
    return (v + np.float32(0.01) * (y * x % np.float32(7.0))).astype(np.float32)[None]


def pyramid_maps(g):
    C = int(g["channels"])
    return [smooth_map(C, int(s), int(s), 0.3 * i) for i, s in enumerate(g["sides"])]


def layer_cases():
    g = load("load_layer2")
    for i in range(int(g["n_cases"])):
        shape = tuple(int(v) for v in g["shape%d" % i])
        ref = np.unpackbits(g["mask_layers%d" % i])[: int(np.prod(shape))].reshape(shape).astype(bool)
        yield g["label%d" % i], int(g["num_classes%d" % i]), ref, g["class_ids%d" % i]


class Cfg:
    RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    IMAGE_SHAPE = np.array([1024, 1024, 3])
    GPU_COUNT = 1
    USE_NMS = True
    DETECTION_MIN_CONFIDENCE = 0.3
    DETECTION_NMS_THRESHOLD = 0.3
    DETECTION_MAX_INSTANCES = 100
    NUM_CLASSES = 3


# ------------------------------------------------------------------ CPU: oracle vs reference Python
def test_oracle_proposal_layer_matches_reference_python():
    g = load("proposal_layer")
    got = oracle.proposal_layer(g["probs"], g["deltas"], g["anchors"], int(g["proposal_count"]), float(g["nms_threshold"]))
    assert got.shape == g["out"][0].shape
    assert got.tobytes() == g["out"][0].tobytes()


def test_oracle_pyramid_roi_align_matches_reference_python():
    g = load("pyramid_roi_align")
    maps = pyramid_maps(g)
    for pool, key in ((7, "pooled7"), (16, "pooled16")):
        got = oracle.pyramid_roi_align(g["boxes"], maps, pool, (256, 256))
        assert got.tobytes() == g[key].tobytes()
    img = oracle.crop_and_resize_fwd(maps[1], g["boxes"], np.zeros(g["boxes"].shape[0], np.int32), 8, 8, 0.0)
    assert img.tobytes() == g["image_crop8"].tobytes()


def test_oracle_refine_detections_matches_reference_python():
    g = load("refine_detections")
    det, keep = oracle.refine_detections(g["rois"], g["probs"], g["deltas"], g["window"])
    assert np.array_equal(keep, g["keep"])
    assert det.tobytes() == g["detections"].tobytes()


def test_oracle_refine_detections_default_branch_matches_reference_python():
    """USE_NMS = False, the reference's shipped default (config.py:78): top-100 by score (Functions.py:526-546)."""
    g = load("refine_detections_nonms")
    for tag in ("a", "b"):
        n = int(g["n_" + tag])
        det, keep = oracle.refine_detections(g["rois"][:n], g["probs"][:n], g["deltas"][:n], g["window"], use_nms=False)
        assert np.array_equal(keep, g["keep_" + tag])
        assert det.tobytes() == g["detections_" + tag].tobytes()
    assert g["detections_a"].shape[0] == 100 and g["detections_b"].shape[0] < 90      # truncated / not truncated


def test_oracle_layer_codec_matches_reference_python():
    for label, nc, ref, class_ids in layer_cases():
        loops = oracle.layer_decode_loops(label, nc)
        assert np.array_equal(loops, ref)
        planes, n_obj = oracle.layer_decode(label, nc - 1, n_max=16)
        assert n_obj == ref.shape[3] == class_ids.size and np.all(class_ids == 1)
        assert np.array_equal(planes[:n_obj].transpose(2, 3, 1, 0).astype(bool), ref)


# ------------------------------------------------------------------ GPU: CUDA path vs reference Python
def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
def test_gpu_proposal_layer_matches_reference_python():
    from sln_amodal_b200 import proposal_layer
    g = load("proposal_layer")
    got = proposal_layer([cuda(g["probs"]).unsqueeze(0), cuda(g["deltas"]).unsqueeze(0)], int(g["proposal_count"]),
                         float(g["nms_threshold"]), cuda(g["anchors"]), Cfg())
    assert tuple(got.shape) == g["out"].shape
    np.testing.assert_allclose(got.cpu().numpy(), g["out"], rtol=3e-7, atol=1e-7)


@pytest.mark.gpu
def test_gpu_pyramid_roi_align_matches_reference_python():
    from sln_amodal_b200 import pyramid_roi_align, pyramid_roi_align_image
    g = load("pyramid_roi_align")
    maps = pyramid_maps(g)
    for cl in (False, True):
        tm = [cuda(m).contiguous(memory_format=torch.channels_last) if cl else cuda(m) for m in maps]
        for pool, key in ((7, "pooled7"), (16, "pooled16")):
            got = pyramid_roi_align([cuda(g["boxes"]).unsqueeze(0)] + tm, pool, (256, 256, 3))
            assert got.contiguous().cpu().numpy().tobytes() == g[key].tobytes()
        img = pyramid_roi_align_image([cuda(g["boxes"]).unsqueeze(0), tm[1]], 8, (256, 256, 3))
        assert img.contiguous().cpu().numpy().tobytes() == g["image_crop8"].tobytes()


@pytest.mark.gpu
def test_gpu_refine_detections_matches_reference_python():
    from sln_amodal_b200 import refine_detections
    g = load("refine_detections")
    det, keep = refine_detections(cuda(g["rois"]), cuda(g["probs"]), cuda(g["deltas"]), g["window"], Cfg())
    assert np.array_equal(keep.cpu().numpy(), g["keep"])
    np.testing.assert_allclose(det.cpu().numpy(), g["detections"], rtol=0, atol=0)


@pytest.mark.gpu
def test_gpu_refine_detections_default_branch_matches_reference_python():
    """The device path of the USE_NMS = False branch (sln_refine_decode + sln_refine_topk) against the reference's own
    refine_detections run with its default config."""
    from sln_amodal_b200 import refine_detections
    g = load("refine_detections_nonms")
    cfg = Cfg()
    cfg.USE_NMS = False
    for tag in ("a", "b"):
        n = int(g["n_" + tag])
        det, keep = refine_detections(cuda(g["rois"][:n]), cuda(g["probs"][:n]), cuda(g["deltas"][:n]), g["window"], cfg)
        assert np.array_equal(keep.cpu().numpy(), g["keep_" + tag])
        assert det.cpu().numpy().tobytes() == g["detections_" + tag].tobytes()


@pytest.mark.gpu
def test_gpu_layer_decode_matches_reference_python():
    from sln_amodal_b200 import decode_layers
    for label, nc, ref, _ in layer_cases():
        planes, n_obj = decode_layers(label, nc, n_max=16)
        n = int(n_obj[0])
        assert n == ref.shape[3]
        got = planes[0, :n].permute(2, 3, 1, 0).cpu().numpy().astype(bool)
        assert np.array_equal(got, ref)
        assert not planes[0, n:].any()


# --------------------------------------------------------------------------- detection targets
def target_cases():
    g = load("detection_targets")
    for k in range(int(g["n_cases"])):
        shp = tuple(int(v) for v in g["masks_shape%d" % k])
        masks = np.unpackbits(g["masks%d" % k])[: int(np.prod(shp))].reshape(shp)
        tshp = tuple(int(v) for v in g["tmasks_shape%d" % k])
        n_t = int(np.prod(tshp)) if len(tshp) > 1 else 0
        tmasks = np.unpackbits(g["tmasks%d" % k])[:n_t].reshape(tshp).astype(np.float32) if n_t else np.zeros(0, np.float32)
        yield dict(props=g["props%d" % k], ids=g["ids%d" % k], gt=g["gt%d" % k], masks=masks, seed=int(g["seed%d" % k]),
                   rois=g["rois%d" % k], cls=g["cls%d" % k], deltas=g["deltas%d" % k], tmasks=tmasks)


class TCfg:
    BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    TRAIN_ROIS_PER_IMAGE = 100
    ROI_POSITIVE_RATIO = 0.7
    MASK_SHAPE = [32, 32]
    USE_MINI_MASK = False
    GPU_COUNT = 1


def test_oracle_detection_targets_match_reference_python():
    g = load("detection_targets")
    assert np.array_equal(oracle.bbox_overlaps(g["ov_b1"], g["ov_b2"]), g["ov_out"], equal_nan=True)
    assert np.array_equal(oracle.box_refinement(g["ref_box"], g["ref_gt"]), g["ref_out"])
    for c in target_cases():
        torch.manual_seed(c["seed"])
        rois, cls, deltas, tm = oracle.detection_target_layer(c["props"], c["ids"], c["gt"], c["masks"])
        if c["rois"].size == 0:
            assert rois.size == 0 and cls.size == 0 and tm.size == 0
            continue
        assert np.array_equal(rois, c["rois"]) and np.array_equal(cls, c["cls"]) and np.array_equal(deltas, c["deltas"])
        assert np.array_equal(tm, c["tmasks"])


@pytest.mark.gpu
def test_gpu_detection_targets_match_reference_python():
    from sln_amodal_b200 import bbox_overlaps, box_refinement, detection_target_layer, ops
    g = load("detection_targets")
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ov = bbox_overlaps(t(g["ov_b1"]), t(g["ov_b2"])).cpu().numpy()
    assert np.array_equal(ov, g["ov_out"], equal_nan=True)
    _, mx, am = ops.bbox_overlaps_device(t(g["ov_b1"]), t(g["ov_b2"]), matrix=False, reduce=True)
    want = torch.from_numpy(g["ov_out"])
    assert np.array_equal(mx.cpu().numpy(), want.max(1)[0].numpy(), equal_nan=True)
    rows = ~np.isnan(g["ov_out"]).any(1)
    assert np.array_equal(am.cpu().numpy()[rows], g["ov_out"][rows].argmax(1))
    np.testing.assert_allclose(box_refinement(t(g["ref_box"]), t(g["ref_gt"])).cpu().numpy(), g["ref_out"], rtol=3e-7, atol=1e-7)
    for c in target_cases():
        torch.manual_seed(c["seed"])
        # GT masks on the host (generic crop path) and on the device as u8 (fused gather + crop + round kernel)
        for gm in (torch.from_numpy(c["masks"]).unsqueeze(0), t(c["masks"]).unsqueeze(0)):
            torch.manual_seed(c["seed"])
            rois, cls, deltas, tm = detection_target_layer(t(c["props"]).unsqueeze(0), t(c["ids"]).unsqueeze(0),
                                                           t(c["gt"]).unsqueeze(0), gm, TCfg())
            if c["rois"].size:
                assert np.array_equal(tm.cpu().numpy(), c["tmasks"]) and np.array_equal(rois.cpu().numpy(), c["rois"])
        if c["rois"].size == 0:
            assert rois.numel() == 0 and cls.numel() == 0 and deltas.numel() == 0 and tm.numel() == 0
            continue
        assert np.array_equal(rois.cpu().numpy(), c["rois"]) and np.array_equal(cls.cpu().numpy(), c["cls"])
        np.testing.assert_allclose(deltas.cpu().numpy(), c["deltas"], rtol=3e-6, atol=1e-6)
        assert np.array_equal(tm.cpu().numpy(), c["tmasks"])


# --------------------------------------------------------------------------- build_rpn_targets
class RCfg:
    RPN_TRAIN_ANCHORS_PER_IMAGE = 256
    RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])


def test_oracle_build_rpn_targets_matches_reference_python():
    g = load("rpn_targets")
    for k in range(int(g["n_cases"])):
        np.random.seed(int(g["rseed%d" % k]))
        m, b = oracle.build_rpn_targets(g["anchors"], g["rids%d" % k], g["rgt%d" % k])
        assert np.array_equal(m, g["rmatch%d" % k]) and np.array_equal(b, g["rbbox%d" % k])


@pytest.mark.gpu
def test_gpu_build_rpn_targets_matches_reference_python():
    from sln_amodal_b200 import build_rpn_targets, ops
    g = load("rpn_targets")
    for k in range(int(g["n_cases"])):
        np.random.seed(int(g["rseed%d" % k]))
        m, b = build_rpn_targets((256, 256, 3), g["anchors"], g["rids%d" % k], g["rgt%d" % k], RCfg())
        assert m.dtype == np.int32 and np.array_equal(m, g["rmatch%d" % k])
        assert b.dtype == np.float64 and np.array_equal(b, g["rbbox%d" % k])
    # the reductions themselves against the float64 matrix, incl. a degenerate (zero-area, 0/0 -> NaN) pair
    anchors = g["anchors"].copy()
    anchors[100] = 0.0
    gt = g["rgt1"].astype(np.float64)
    gt[0] = 0.0
    ov = oracle.compute_overlaps(anchors, gt)
    dev = torch.device("cuda", 0)
    mx, am, ga = ops.rpn_overlap_reductions_device(torch.from_numpy(anchors).to(dev), torch.from_numpy(gt))
    assert np.array_equal(am.cpu().numpy(), np.argmax(ov, axis=1))
    assert np.array_equal(mx.cpu().numpy(), ov[np.arange(ov.shape[0]), np.argmax(ov, axis=1)], equal_nan=True)
    assert np.array_equal(ga.cpu().numpy(), np.argmax(ov, axis=0))


def test_rpn_pack_oracle_matches_reference_module():
    """oracle.rpn_pack against the reference's own RPN module + MaskRCNN.predict concatenation (fixture made by
    tests/golden/make_golden_rpn.py): copies bit for bit, softmax within 2 ulp of torch's CPU kernel (numpy's exp)."""
    g = np.load(os.path.join(G, "rpn_pack.npz"))
    logits, probs, bbox = oracle.rpn_pack([g["cls_%d" % l] for l in range(3)], [g["box_%d" % l] for l in range(3)])
    assert np.array_equal(logits, g["rpn_class_logits"])
    assert np.array_equal(bbox, g["rpn_bbox"])
    assert np.allclose(probs, g["rpn_class"], rtol=3e-7, atol=0)


def test_rpn_pack_oracle_matches_torch_expression_random_shapes():
    """oracle.rpn_pack against the reference's expression (modal/modals.py:394-410 + model.py:553-563) evaluated with
    torch on the CPU, over random level counts, sizes, batch sizes and anchors per location."""
    rng = np.random.default_rng(17)
    for _ in range(12):
        a, B, nl = int(rng.integers(1, 5)), int(rng.integers(1, 4)), int(rng.integers(1, 6))
        shapes = [(int(rng.integers(1, 9)), int(rng.integers(1, 9))) for _ in range(nl)]
        cm = [rng.standard_normal((B, 2 * a, h, w)).astype(np.float32) * 5 for h, w in shapes]
        bm = [rng.standard_normal((B, 4 * a, h, w)).astype(np.float32) for h, w in shapes]
        lg = [torch.from_numpy(c).permute(0, 2, 3, 1).contiguous().view(B, -1, 2) for c in cm]
        bx = [torch.from_numpy(b).permute(0, 2, 3, 1).contiguous().view(B, -1, 4) for b in bm]
        want = (torch.cat(lg, 1).numpy(), torch.cat([torch.softmax(x, dim=2) for x in lg], 1).numpy(), torch.cat(bx, 1).numpy())
        got = oracle.rpn_pack(cm, bm)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2])
        assert np.allclose(got[1], want[1], rtol=3e-7, atol=0)


def test_unmold_oracle_matches_reference_unmold_detections():
    """oracle.unmold_mask against the masks the reference's own MaskRCNN.unmold_detections / utils.unmold_mask produced
    (tests/golden/make_golden_unmold.py: reference Python unmodified, scipy.misc.imresize supplied around the real Pillow)."""
    g = np.load(os.path.join(G, "unmold_detections.npz"))
    boxes, masks = g["boxes"], g["masks"]
    kept = [i for i in range(7) if i != 2]                          # detection 2 has zero area and is dropped
    assert masks.shape == (96, 128, 6) and boxes.shape == (6, 4)
    for j, i in enumerate(kept):
        want = masks[:, :, j]
        got = oracle.unmold_mask(g["mrcnn_mask"][i, :, :, 1], boxes[j], tuple(g["image_shape"]))
        assert np.array_equal(got, want), (i, boxes[j].tolist())


def test_unmold_detections_host_logic_matches_reference(monkeypatch):
    """The host half of sln_amodal_b200.unmold.unmold_detections (padding cut, class collapse, window scale / shift,
    int32 cast, zero-area filter, output layout) against the reference's own function, with the device launch replaced by
    the oracle so that the test runs without a GPU."""
    from sln_amodal_b200 import unmold
    g = np.load(os.path.join(G, "unmold_detections.npz"))

    def fake_unmold_masks(masks, boxes, image_shape):
        m = masks.numpy() if isinstance(masks, torch.Tensor) else np.asarray(masks)
        return torch.from_numpy(np.stack([oracle.unmold_mask(m[i], boxes[i], image_shape) for i in range(m.shape[0])]))
    monkeypatch.setattr(unmold, "unmold_masks", fake_unmold_masks)
    boxes, class_ids, scores, masks = unmold.unmold_detections(g["detections"], g["mrcnn_mask"], tuple(g["image_shape"]), g["window"])
    assert boxes.dtype == g["boxes"].dtype and np.array_equal(boxes, g["boxes"])
    assert np.array_equal(class_ids, g["class_ids"]) and np.array_equal(scores, g["scores"])
    assert masks.shape == g["masks"].shape and np.array_equal(masks, g["masks"])


def test_resize_image_oracle_matches_reference_resize_image():
    """oracle.resize_image against the reference's own utils.resize_image output (same generator)."""
    g = np.load(os.path.join(G, "resize_image.npz"))
    assert np.array_equal(oracle.resize_image(g["image"], (64, 64)), g["resized"])
    assert tuple(g["window"]) == (0, 0, 64, 64) and np.allclose(g["scale"], (64 / 75, 64 / 50))


# --------------------------------------------------------------------------- load_image_gt (section 8(f)-2 glue)
class _FakeDataset:
    """the two members load_image_gt touches (modal/Functions.py:697-699): image_info[...]['path'] and load_image"""

    def __init__(self, path, image):
        self.image_info = [{"path": path, "height": image.shape[0], "width": image.shape[1], "id": 0}]
        self._image = image

    def load_image(self, image_id):
        return self._image


class _GtCfg:
    IMAGE_PADDING = True
    USE_MINI_MASK = False
    MINI_MASK_SHAPE = (56, 56)


def _loadgt_cases():
    g = np.load(os.path.join(G, "load_image_gt.npz"))
    for i in range(int(g["n_cases"])):
        shape = tuple(g["masks_shape%d" % i])
        masks = np.unpackbits(g["masks%d" % i])[: int(np.prod(shape))].reshape(shape)
        yield i, g, masks


def test_oracle_load_image_gt_matches_reference():
    """the numpy restatement of load_image_gt reproduces what the reference's own function returned (same seeds)"""
    for i, g, masks in _loadgt_cases():
        img, meta, cls, bbox, m = oracle.load_image_gt(g["label%d" % i], g["image_in%d" % i], int(g["num_classes%d" % i]),
                                                       int(g["max_dim"]), bool(g["augment%d" % i]), int(g["seed%d" % i]))
        assert np.array_equal(img, g["image%d" % i]), i
        assert np.array_equal(meta, g["meta%d" % i]) and np.array_equal(cls, g["class_ids%d" % i]), i
        assert np.array_equal(bbox, g["bbox%d" % i]), i
        assert m.dtype == np.uint8 and np.array_equal(m, masks), i


@pytest.mark.gpu
def test_load_image_gt_device_path_matches_reference(tmp_path):
    """targets.load_image_gt (npz reader -> layer decode -> Pillow-exact resize -> zoom gather + flip -> plane boxes) against
    the reference's own load_image_gt output, value for value, with the generators seeded like the reference run"""
    import random
    from sln_amodal_b200 import targets
    for i, g, masks in _loadgt_cases():
        path = str(tmp_path / ("img%d.jpg" % i))
        np.savez_compressed(path[:-4] + ".npz", layer=g["label%d" % i])
        ds = _FakeDataset(path, g["image_in%d" % i])
        cfg = _GtCfg()
        cfg.NUM_CLASSES = int(g["num_classes%d" % i])
        cfg.IMAGE_MAX_DIM, cfg.IMAGE_MIN_DIM = int(g["max_dim"]), int(g["min_dim"])
        random.seed(int(g["seed%d" % i]))
        np.random.seed(int(g["seed%d" % i]))
        img, meta, cls, bbox, m = targets.load_image_gt(ds, cfg, 0, augment=bool(g["augment%d" % i]))
        assert isinstance(img, np.ndarray) and np.array_equal(img, g["image%d" % i]), i
        assert np.array_equal(meta, g["meta%d" % i]) and np.array_equal(cls, g["class_ids%d" % i]), i
        assert np.array_equal(bbox, g["bbox%d" % i]), i
        assert m.dtype == np.uint8 and m.shape == masks.shape and np.array_equal(m, masks), i
        # device=True: the same planes, left on the device in [n, L, H, W]
        random.seed(int(g["seed%d" % i]))
        np.random.seed(int(g["seed%d" % i]))
        planes = targets.load_image_gt(ds, cfg, 0, augment=bool(g["augment%d" % i]), device=True)[4]
        assert planes.is_cuda and np.array_equal(planes.permute(2, 3, 0, 1).cpu().numpy(), masks)
    with pytest.raises(ValueError):
        targets.load_image_gt(ds, cfg, 0, use_mini_mask=True)

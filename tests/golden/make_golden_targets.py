"""Golden vectors for the SURVEY 8(f)-1 row: bbox_overlaps, utils.box_refinement and detection_target_layer
(modal/Functions.py:184-416, utils.py:96-117), produced by the REFERENCE's own Python, unmodified, imported from
/root/reference (same shims and stubs as make_golden.py).  Run in the build container only:

    python tests/golden/make_golden_targets.py

detection_target_layer samples with torch.randperm on the CPU generator: the fixture stores the seed, and the mirror
in this repo draws from the same generator in the same order, so outputs can be compared element for element.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Cfg, setup_reference_imports  # noqa: E402


class TCfg(Cfg):
    TRAIN_ROIS_PER_IMAGE = 100
    ROI_POSITIVE_RATIO = 0.7
    MASK_SHAPE = [32, 32]
    USE_MINI_MASK = False


def make_inputs(seed, n_prop=400, n_gt=7, L=2, side=128, crowds=0):
    rng = np.random.default_rng(seed)
    gt = np.zeros((n_gt, 4), np.float32)
    masks = np.zeros((L, n_gt, side, side), np.uint8)
    for i in range(n_gt):
        h, w = rng.uniform(0.12, 0.45, 2)
        y1, x1 = rng.uniform(0, 1 - h), rng.uniform(0, 1 - w)
        gt[i] = (y1, x1, y1 + h, x1 + w)
        ys, xs = int(y1 * side), int(x1 * side)
        ye, xe = max(ys + 2, int((y1 + h) * side)), max(xs + 2, int((x1 + w) * side))
        yy, xx = np.mgrid[ys:ye, xs:xe]
        ell = ((yy - (ys + ye) / 2) / ((ye - ys) / 2)) ** 2 + ((xx - (xs + xe) / 2) / ((xe - xs) / 2)) ** 2 <= 1.0
        masks[0, i, ys:ye, xs:xe] = ell
        masks[1, i, ys:ye, xs:xe] = ell & (xx > (xs + xe) / 2)          # an "occluded half" layer
    # proposals: jittered GT boxes (positives), random boxes (negatives), some zero padding
    reps = rng.integers(0, n_gt, n_prop // 2)
    jit = gt[reps] + rng.normal(0, 0.03, (n_prop // 2, 4)).astype(np.float32)
    rnd_c = rng.uniform(0.1, 0.9, (n_prop // 2 - 10, 2))
    rnd_s = rng.uniform(0.03, 0.3, (n_prop // 2 - 10, 2))
    rnd = np.concatenate([rnd_c - rnd_s / 2, rnd_c + rnd_s / 2], 1)
    props = np.clip(np.concatenate([jit, rnd, np.zeros((10, 4))], 0), 0, 1).astype(np.float32)
    ids = rng.integers(1, 3, n_gt).astype(np.int32)
    if crowds:
        ids[:crowds] = -1
    return props, ids, gt, masks


def main():
    setup_reference_imports()
    import modal.Functions as F
    import utils as U
    out = {}
    cases = {}
    # ---- bbox_overlaps / box_refinement on their own
    rng = np.random.default_rng(7)
    b1 = np.clip(rng.uniform(0, 1, (300, 2)).repeat(2, 1)[:, [0, 2, 1, 3]] + np.array([0, 0, 0.2, 0.3]) * rng.uniform(0.1, 1, (300, 4)), 0, 1).astype(np.float32)
    b2 = np.clip(rng.uniform(0, 1, (9, 2)).repeat(2, 1)[:, [0, 2, 1, 3]] + np.array([0, 0, 0.3, 0.2]) * rng.uniform(0.1, 1, (9, 4)), 0, 1).astype(np.float32)
    b1[5] = 0.0                                                     # zero-padded proposal
    ov = F.bbox_overlaps(torch.from_numpy(b1), torch.from_numpy(b2)).numpy()
    cases.update(ov_b1=b1, ov_b2=b2, ov_out=ov, ref_box=b1[10:110])
    gt_sel = b2[np.random.default_rng(70).integers(0, 9, 100)]
    ref = U.box_refinement(torch.from_numpy(b1[10:110]), torch.from_numpy(gt_sel)).numpy()
    cases.update(ref_gt=gt_sel, ref_out=ref)
    # ---- the whole layer, three inputs (plain, with crowd boxes, no positives)
    for k, (seed, crowds, shift) in enumerate([(11, 0, 0.0), (12, 2, 0.0), (13, 0, 5.0)]):
        props, ids, gt, masks = make_inputs(seed, crowds=crowds)
        if shift:
            gt = (gt + shift).astype(np.float32)                     # GT far away: no positive ROI
        torch.manual_seed(1000 + k)
        rois, cls, deltas, m = F.detection_target_layer(torch.from_numpy(props).unsqueeze(0), torch.from_numpy(ids).unsqueeze(0),
                                                        torch.from_numpy(gt).unsqueeze(0), torch.from_numpy(masks).unsqueeze(0), TCfg())
        cases.update({"props%d" % k: props, "ids%d" % k: ids, "gt%d" % k: gt, "masks%d" % k: np.packbits(masks),
                      "masks_shape%d" % k: np.array(masks.shape), "seed%d" % k: 1000 + k,
                      "rois%d" % k: rois.numpy(), "cls%d" % k: np.asarray(cls.numpy()), "deltas%d" % k: deltas.numpy(),
                      "tmasks%d" % k: np.packbits(m.numpy().astype(np.uint8)) if m.numel() else np.zeros(0, np.uint8),
                      "tmasks_shape%d" % k: np.array(m.shape)})
        out["case%d" % k] = (tuple(rois.shape), tuple(m.shape), int((cls.numpy() > 0).sum()) if cls.numel() else 0)
    np.savez_compressed(os.path.join(HERE, "detection_targets.npz"), n_cases=3, **cases)
    print("written:", out, "overlaps", ov.shape, "refinement", ref.shape)


def rpn_targets_main():
    """build_rpn_targets (modal/Functions.py:739-847) on a reduced anchor pyramid: golden rpn_match / rpn_bbox."""
    setup_reference_imports()
    import modal.Functions as F
    import utils as U

    class RCfg(Cfg):
        RPN_TRAIN_ANCHORS_PER_IMAGE = 256

    anchors = U.generate_pyramid_anchors((32, 64, 128, 256, 512), [0.5, 1, 2], [[64, 64], [32, 32], [16, 16], [8, 8], [4, 4]],
                                         [4, 8, 16, 32, 64], 1)          # 256^2 image: 16368 anchors
    cases = {"anchors": anchors}
    out = {}
    for k, (seed, n_gt, crowds) in enumerate([(21, 6, 0), (22, 9, 2), (23, 1, 0)]):
        rng = np.random.default_rng(seed)
        c = rng.uniform(30, 226, (n_gt, 2))
        s = rng.uniform(12, 110, (n_gt, 2))
        gt = np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 256).astype(np.int32)
        ids = rng.integers(1, 3, n_gt).astype(np.int32)
        if crowds:
            ids[:crowds] = -1
        np.random.seed(500 + k)
        match, bbox = F.build_rpn_targets((256, 256, 3), anchors, ids, gt, RCfg())
        cases.update({"rgt%d" % k: gt, "rids%d" % k: ids, "rseed%d" % k: 500 + k, "rmatch%d" % k: match, "rbbox%d" % k: bbox})
        out[k] = (int((match == 1).sum()), int((match == -1).sum()))
    np.savez_compressed(os.path.join(HERE, "rpn_targets.npz"), n_cases=3, **cases)
    print("rpn targets written:", out, anchors.shape, anchors.dtype)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rpn":
        rpn_targets_main()
    else:
        main()

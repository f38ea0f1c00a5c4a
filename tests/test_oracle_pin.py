"""Pins the oracle restatement (oracle/sln_oracle.c + oracle/oracle.py) before anything trusts it:
  * bit-for-bit against the reference's own unmodified C (oracle/_ref, built from
    /root/reference/roialign/roi_align/src/crop_and_resize.c and /root/reference/nms/src/nms.c);
  * the closed-form layer decode against the loop-for-loop restatement of load_layer2;
  * the EDT against scipy.ndimage.distance_transform_edt (the reference has no EDT).
The reference ships no golden vectors for this path, so its compiled
sources are the pin."""
import numpy as np
import pytest

from oracle import oracle
from sln_amodal_b200 import synth

needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")


def _crop_case(seed, B, C, H, W, N, outside=0.1, degenerate=0.05):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((B, C, H, W), dtype=np.float32)
    boxes = synth.roi_boxes(N, seed=seed + 1, outside_frac=outside, degenerate_frac=degenerate)
    ind = rng.integers(0, B, N).astype(np.int32)
    return img, boxes, ind


@needs_ref
@pytest.mark.parametrize("ph,pw", [(7, 7), (14, 14), (16, 16), (1, 1), (1, 5), (3, 2)])
@pytest.mark.parametrize("C", [1, 3, 8])
def test_crop_fwd_matches_reference_c(ph, pw, C):
    img, boxes, ind = _crop_case(11 + ph * 3 + C, 2, C, 33, 47, 64)
    for ext in (0.0, -2.5):
        a = oracle.crop_and_resize_fwd(img, boxes, ind, ph, pw, ext)
        b = oracle.ref_crop_and_resize_fwd(img, boxes, ind, ph, pw, ext)
        assert a.tobytes() == b.tobytes()


@needs_ref
@pytest.mark.parametrize("ph,pw", [(7, 7), (14, 14), (1, 1), (2, 5)])
def test_crop_bwd_matches_reference_c(ph, pw):
    img, boxes, ind = _crop_case(5 + ph, 3, 4, 29, 31, 80)
    rng = np.random.default_rng(99)
    g = rng.standard_normal((boxes.shape[0], 4, ph, pw), dtype=np.float32)
    a = oracle.crop_and_resize_bwd(g, boxes, ind, img.shape)
    b = oracle.ref_crop_and_resize_bwd(g, boxes, ind, img.shape)
    assert a.tobytes() == b.tobytes()


def test_crop_integral_and_edge_samples():
    # boxes landing exactly on pixel centres: floor == ceil, lerp == 0
    img = np.arange(2 * 1 * 5 * 5, dtype=np.float32).reshape(2, 1, 5, 5)
    boxes = np.array([[0, 0, 1, 1], [0.25, 0.25, 0.75, 0.75], [1, 1, 1, 1], [-0.5, 0, 0.5, 1]], np.float32)
    ind = np.array([0, 1, 1, 0], np.int32)
    out = oracle.crop_and_resize_fwd(img, boxes, ind, 5, 5, -1.0)
    assert np.array_equal(out[0, 0], img[0, 0])
    assert np.array_equal(out[1, 0, ::2, ::2], img[1, 0, 1:4, 1:4])
    assert np.all(out[2] == img[1, 0, 4, 4])
    assert np.all(out[3, 0, :2] == -1.0) and np.array_equal(out[3, 0, 2], img[0, 0, 0])
    with pytest.raises(ValueError):
        oracle.crop_and_resize_fwd(img, boxes, np.array([0, 2, 0, 0], np.int32), 2, 2)


@needs_ref
@pytest.mark.parametrize("kind", ["rpn", "uniform"])
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 500, 1500])
@pytest.mark.parametrize("ties", [False, True])
def test_nms_matches_reference_c(kind, n, ties):
    boxes = synth.nms_boxes(n, seed=7 + n, kind=kind)
    scores = synth.nms_scores(n, seed=8 + n, ties=ties)
    dets = np.concatenate([boxes, scores[:, None]], 1)
    for thr in (0.7, 0.3):
        a = oracle.nms(dets, thr)
        b = oracle.ref_nms(dets, thr)
        assert np.array_equal(a, b)
        assert a.size >= 1


def test_nms_rounded_boxes_and_empty():
    assert oracle.nms(np.zeros((0, 5), np.float32), 0.5).size == 0
    boxes = synth.nms_boxes(300, seed=3, rounded=True)
    dets = np.concatenate([boxes, synth.nms_scores(300)[:, None]], 1)
    k = oracle.nms(dets, 0.3)
    assert len(set(k.tolist())) == k.size
    # kept boxes are pairwise below threshold under the +1 convention
    kb = boxes[k]
    for i in range(min(20, k.size)):
        yy1 = np.maximum(kb[i, 0], kb[:, 0]); xx1 = np.maximum(kb[i, 1], kb[:, 1])
        yy2 = np.minimum(kb[i, 2], kb[:, 2]); xx2 = np.minimum(kb[i, 3], kb[:, 3])
        inter = np.maximum(0, yy2 - yy1 + 1) * np.maximum(0, xx2 - xx1 + 1)
        area = (kb[:, 2] - kb[:, 0] + 1) * (kb[:, 3] - kb[:, 1] + 1)
        iou = inter / (area[i] + area - inter)
        iou[i] = 0
        assert np.all(iou < 0.3)


@pytest.mark.parametrize("num_classes", [2, 3, 4, 6])
def test_layer_closed_form_matches_loops(num_classes):
    label = synth.label_map(96, 128, n=9, seed=2024 + num_classes, min_piece=16)
    L = num_classes - 1
    ref = oracle.layer_decode_loops(label, num_classes)     # bool [H,W,L,n]
    out, n_obj = oracle.layer_decode(label, L, n_max=32)
    assert ref is not None and ref.shape[3] == n_obj
    assert np.array_equal(out[:n_obj].transpose(2, 3, 1, 0).astype(bool), ref)
    assert not out[n_obj:].any()


def test_layer_edge_labels():
    # overlapping annotations: a label with two visible bits, a label that is occluded-only,
    # and an object that is never the top visible bit (truncates n_obj, Functions.py:1074-1079)
    label = np.zeros((8, 8), np.uint64)
    label[0, :4] = (1 << 0) | (1 << 1)              # objects 0 and 1 both "visible"
    label[1, :4] = (1 << 0)
    label[2, :4] = (1 << 0) | (1 << (32 + 1)) | (1 << (32 + 3))
    label[3, :4] = (1 << (32 + 2))                  # occluded-only piece
    label[4, :4] = (1 << 3)                         # object 3 visible, object 2 never visible
    for nc in (2, 3, 5):
        ref = oracle.layer_decode_loops(label, nc)
        out, n_obj = oracle.layer_decode(label, nc - 1, n_max=8)
        assert n_obj == 2 and ref.shape[3] == 2
        assert np.array_equal(out[:n_obj].transpose(2, 3, 1, 0).astype(bool), ref)
    assert oracle.layer_decode_loops(np.zeros((4, 4), np.uint64), 3) is None
    assert oracle.layer_decode(np.zeros((4, 4), np.uint64), 2, n_max=4)[1] == 0


def test_edt_matches_scipy():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(5)
    cases = []
    for H, W, p in [(1, 1, 0.5), (1, 17, 0.7), (19, 1, 0.7), (37, 53, 0.9), (64, 64, 0.995), (128, 96, 0.5)]:
        cases.append((rng.random((H, W)) < p).astype(np.uint8))
    lab = synth.label_map(160, 200, n=6, seed=9, min_piece=16)
    cases.append(((lab & np.uint64(1)) != 0).astype(np.uint8))
    cases.append(((lab >> np.uint64(33)) & np.uint64(1)).astype(np.uint8))
    for m in cases:
        if m.all():
            m.flat[0] = 0
        want = np.rint(ndi.distance_transform_edt(m) ** 2).astype(np.int64)
        got = oracle.edt_sq(m)
        assert np.array_equal(got.astype(np.int64), want)
    full = np.ones((5, 7), np.uint8)
    assert np.all(oracle.edt_sq(full) == (5 + 7) ** 2)


def test_rle_restatement_matches_reference_maskapi():
    """oracle.rle_encode / rle_to_string against the reference's own maskApi.c (oracle/_ref/libref_mask.so)."""
    import pytest
    if not oracle.ref_mask_available():
        pytest.skip("oracle/_ref/libref_mask.so not built (no /root/reference here)")
    rng = np.random.default_rng(1)
    h, w = 41, 29
    masks = (rng.random((6, h * w)) < 0.35).astype(np.uint8)
    masks[1] = 0
    masks[2] = 1
    masks[3, :7] = 1
    masks[4, -1] = 1
    ref = oracle.ref_rle_encode(masks, h, w)
    for i in range(6):
        c = oracle.rle_encode(masks[i])
        assert np.array_equal(c, ref[i][0]) and oracle.rle_to_string(c) == ref[i][1]


def test_unmold_restatement_matches_pillow():
    """oracle.pil_resize_bilinear_u8 / unmold_mask (the restatement the CUDA kernel follows) against the real Pillow --
    the library scipy.misc.imresize(interp='bilinear'), called by utils.unmold_mask (utils.py:459-460), runs."""
    import pytest
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for t in range(120):
        mh, mw = (28, 28) if t % 3 else (int(rng.integers(1, 40)), int(rng.integers(1, 40)))
        img = rng.integers(0, 256, (mh, mw)).astype(np.uint8)
        oh, ow = int(rng.integers(1, 160)), int(rng.integers(1, 160))
        want = np.asarray(Image.frombytes("L", (mw, mh), img.tobytes()).resize((ow, oh), resample=Image.BILINEAR))
        assert np.array_equal(oracle.pil_resize_bilinear_u8(img, oh, ow), want), (mh, mw, oh, ow)
    for t in range(40):
        m = rng.random((28, 28)).astype(np.float32) ** 2
        if t == 0:
            m[:] = 0.25                                                     # constant mask: cscale == 0 branch
        y1, x1 = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        y2, x2 = y1 + int(rng.integers(0, 70)), x1 + int(rng.integers(0, 70))
        a = oracle.unmold_mask(m, (y1, x1, y2, x2), (128, 130, 3))
        b = oracle.unmold_mask_pil(m, (y1, x1, y2, x2), (128, 130, 3))
        assert a.shape == (128, 130) and np.array_equal(a, b)


def test_bytescale_restatement_is_float32():
    """scipy.misc.bytescale under numpy-1.x promotion computes in float32 (the array's dtype); the restatement must not
    depend on the numpy-2 promotion rules of the interpreter it runs on."""
    m = np.linspace(0, 1, 784, dtype=np.float32).reshape(28, 28) ** 3
    b = oracle.bytescale_f32(m)
    assert b.dtype == np.uint8 and b.min() == 0 and b.max() == 255
    lo, hi = m.min(), m.max()
    scale = np.float32(255.0 / float(np.float32(hi - lo)))
    want = (np.clip((m - lo) * scale, 0, 255).astype(np.float32) + np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(b, want)


def test_resize_image_restatement_matches_pillow_rgb():
    """oracle.resize_image (per-band restatement) against the real Pillow on RGB images -- what utils.resize_image's
    scipy.misc.imresize(image, (max_dim, max_dim)) runs (utils.py:352): up- and down-scaling, odd sizes."""
    import pytest
    pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(9)
    for (h, w, H2, W2) in ((37, 53, 64, 64), (120, 90, 64, 64), (64, 64, 64, 64), (200, 31, 48, 80), (5, 7, 33, 2)):
        img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        assert np.array_equal(oracle.resize_image(img, (H2, W2)), oracle.resize_image_pil(img, (H2, W2))), (h, w, H2, W2)
    g = rng.integers(0, 256, (40, 50)).astype(np.uint8)
    assert np.array_equal(oracle.resize_image(g, (64, 64)), oracle.resize_image_pil(g, (64, 64)))


def test_banded_edt_restatement_equals_the_first_one_and_scipy():
    """orc_edt_sq_banded (row pass + per-band lower envelopes + minimum over the bands in reach: the column algorithm
    planned for the CUDA kernel) against orc_edt_sq (column pass + bounded search) and scipy, for band heights from 1 to
    more than H, on blobs, noise, single zeros and maps without a zero."""
    from scipy import ndimage as ndi
    from sln_amodal_b200 import synth
    rng = np.random.default_rng(41)
    lab = synth.label_map(96, 112, n=5, seed=9, min_piece=16)
    cases = [((lab & np.uint64(1)) != 0).astype(np.uint8), ((lab >> np.uint64(33)) & np.uint64(1)).astype(np.uint8),
             (rng.random((70, 45)) < 0.93).astype(np.uint8), (rng.random((33, 64)) < 0.5).astype(np.uint8)]
    one = np.ones((50, 37), np.uint8)
    one[17, 30] = 0
    cases.append(one)
    cases.append(np.ones((9, 13), np.uint8))                       # no zero pixel: (H + W)^2 everywhere
    cases.append(np.zeros((8, 8), np.uint8))
    for m in cases:
        want = oracle.edt_sq(m)
        if not m.all():
            assert np.array_equal(want.astype(np.int64), np.rint(ndi.distance_transform_edt(m) ** 2).astype(np.int64))
        for band in (1, 2, 5, 16, 32, 64, 1000):
            assert np.array_equal(oracle.edt_sq_banded(m, band), want), (m.shape, band)

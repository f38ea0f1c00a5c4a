"""CPU-side tests: the C-ABI library loads and exports every symbol include/sln_b200.h declares
(no compute calls without a GPU), the host logic of the operator mirror, the import-path shadows,
and the multi-process sharding / gather logic over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from sln_amodal_b200 import build
    return build.build()


def test_header_symbols_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "sln_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sln_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 17
    handle = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in sln_b200.h but not exported"
    from sln_amodal_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared, "ctypes prototypes and header disagree"
    assert _lib.lib().sln_version() >= 100
    assert _lib.lib().sln_last_error_string() is not None


def test_header_flag_values_match_the_python_constants():
    """The flag / layout macros of include/sln_b200.h and the constants the ctypes layer passes are the same numbers
    (SLN_BWD_PLAN_ONLY / SLN_BWD_PLANNED: the backward planned beside the forward)."""
    from sln_amodal_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sln_b200.h")).read()
    macros = {k: int(v) for k, v in re.findall(r"#define\s+(SLN_[A-Z0-9_]+)\s+(-?\d+)\b", hdr)}
    assert macros["SLN_BWD_EXACT"] == _lib.BWD_EXACT
    assert macros["SLN_BWD_PLAN_ONLY"] == _lib.BWD_PLAN_ONLY
    assert macros["SLN_BWD_PLANNED"] == _lib.BWD_PLANNED
    assert len({macros["SLN_BWD_EXACT"], macros["SLN_BWD_PLAN_ONLY"], macros["SLN_BWD_PLANNED"]}) == 3     # distinct bits
    assert macros["SLN_BWD_EXACT"] & macros["SLN_BWD_PLAN_ONLY"] == 0 and macros["SLN_BWD_PLAN_ONLY"] & macros["SLN_BWD_PLANNED"] == 0


def test_workspace_queries_and_arg_errors(built_lib):
    """Pure host-side entry points: sizes and argument validation (no kernel is launched)."""
    from sln_amodal_b200 import _lib
    L = _lib.lib()
    assert L.sln_nms_workspace_bytes(0) >= 0
    n = 12000
    assert L.sln_nms_workspace_bytes(n) >= n * ((n + 63) // 64) * 8
    assert L.sln_crop_and_resize_bwd_workspace_bytes(8000, 8, 7, 7) >= 8000 * 12
    # banded path: 34 stack slots per 32 rows (u32) + two words per (band, column) + tile flags
    assert L.sln_edt_workspace_bytes(320, 1024, 1024) <= 320 * 1024 * 1024 * 4 * (34 / 32 + 2 / 32) + (2 << 20)
    assert L.sln_edt_workspace_bytes(4, 4096, 4096) <= 4 * 4096 * 4096 * 2 + (2 << 20)
    assert L.sln_proposal_workspace_bytes(261888, 6000) == L.sln_proposal_workspace_bytes(10 ** 6, 6000)
    # argument validation happens before any CUDA call
    rc = L.sln_crop_and_resize_fwd(None, 1, 1, 4, 4, 0, None, None, 1, 0, 7, 0.0, None, None)
    assert rc == -1 and b"crop size" in L.sln_last_error_string()
    rc = L.sln_crop_and_resize_bwd(None, None, None, 1, 1, 7, 7, None, 1, 4, 4, 0, 0, None, 0, None)
    assert rc == -2 and b"NHWC" in L.sln_last_error_string()
    rc = L.sln_nms(None, None, -1, 0.5, 0, None, None, None, 0, None)
    assert rc == -1
    rc = L.sln_layer_decode(None, 1, 4, 4, 0, 4, None, None, None, None)
    assert rc == -1


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: CPU tensors raise instead of silently running somewhere else."""
    from sln_amodal_b200 import CropAndResizeFunction, nms, _lib
    with pytest.raises(_lib.SlnError):
        CropAndResizeFunction(7, 7, 0)(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4), torch.zeros(1).int())
    with pytest.raises(_lib.SlnError):
        nms(torch.zeros(4, 5), 0.5)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under sln_amodal_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sln_amodal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/_ref" not in src and "libsln_oracle" not in src, f


def test_import_path_shadows():
    import roialign.roi_align.crop_and_resize as a
    import roialign.roi_align.roi_align as b
    import nms.nms_wrapper as c
    import nms.pth_nms as d
    import sln_amodal_b200 as s
    assert a.CropAndResizeFunction is s.CropAndResizeFunction and a.CropAndResize is s.CropAndResize
    assert b.RoIAlign is s.RoIAlign
    assert c.nms is s.nms and d.pth_nms is s.pth_nms
    f = a.CropAndResizeFunction(7, 14, 0.5)
    assert (f.crop_height, f.crop_width, f.extrapolation_value) == (7, 14, 0.5)


def test_install_rebinds_named_functions():
    import types
    import sln_amodal_b200 as s
    model = types.ModuleType("model")
    fn = types.ModuleType("modal.Functions")
    md = types.ModuleType("modal.modals")
    for m in (model, fn):
        m.proposal_layer = m.refine_detections = m.pyramid_roi_align_image = lambda *a, **k: None
    md.pyramid_roi_align = md.pyramid_roi_align_image = lambda *a, **k: None

    class DS:
        pass
    done = s.install(model, fn, md, dataset_class=DS)
    assert model.proposal_layer is s.proposal_layer and fn.proposal_layer is s.proposal_layer
    assert md.pyramid_roi_align is s.pyramid_roi_align
    assert model.pyramid_roi_align_image is s.pyramid_roi_align_image
    assert DS.load_layer2 is s.load_layer2
    assert "model.proposal_layer" in done and "DS.load_layer2" in done


def test_roi_level_matches_oracle_on_cpu():
    """pyramid.roi_level is plain torch; check it against the oracle's restatement of modals.py:53-64."""
    from oracle import oracle
    from sln_amodal_b200 import synth
    from sln_amodal_b200.pyramid import roi_level
    boxes = synth.roi_boxes(5000, seed=3)
    got = roi_level(torch.from_numpy(boxes), (1024, 1024, 3)).numpy()
    assert np.array_equal(got, oracle.roi_levels(boxes))
    hist = np.bincount(got, minlength=6)[2:]
    assert hist.sum() == 5000 and (hist > 0).all()


def test_synth_generators_are_deterministic():
    from sln_amodal_b200 import synth
    a, b = synth.roi_boxes(100, seed=4321), synth.roi_boxes(100, seed=4321)
    assert a.tobytes() == b.tobytes() and a.dtype == np.float32 and (a >= 0).all() and (a <= 1).all()
    lab = synth.label_map(64, 64, n=5, seed=1, min_piece=8)
    assert lab.dtype == np.uint64 and lab.any()
    # writer rule: a pixel has at most one visible owner; occluded bits only under a visible one
    lo = lab & np.uint64(0xFFFFFFFF)
    assert np.all((lo & (lo - np.uint64(1))) == 0)
    s = synth.nms_scores(1000, seed=8)
    assert np.unique(s).size == 1000
    assert np.unique(synth.nms_scores(1000, seed=8, ties=True)).size <= 256


def test_shard_range_partitions():
    from sln_amodal_b200.dist import shard_range
    for n in (0, 1, 7, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["SLN_ROOT"])
import torch.distributed as dist
from sln_amodal_b200 import dist as sdist
rank, world = sdist.init(backend="gloo")
assert world == 2
n_images = 5
lo, hi = sdist.shard_range(n_images, rank, world)
torch.manual_seed(0)
box_ind = torch.randint(0, n_images, (200,))
boxes = torch.rand(200, 4)
b, ind, span = sdist.shard_rois(boxes, box_ind, n_images, rank, world)
assert span == (lo, hi) and ((ind >= 0) & (ind < hi - lo)).all()
total = sdist.sum_over_ranks(float(b.shape[0]))
assert total == 200.0, total
# per-image "detections": image i has i+1 rows filled with i
local = [torch.full((i + 1, 6), float(i)) for i in range(lo, hi)]
allv = sdist.gather_detections(local, max_per_image=10, width=6)
assert len(allv) == n_images
for i, d in enumerate(allv):
    assert d.shape == (i + 1, 6) and (d == i).all()
assert sdist.max_over_ranks(float(rank)) == 1.0
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, SLN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2


def test_zoom_index_map_reproduces_scipy_nearest_zoom():
    """targets.zoom_index_map is the host half of resize_layer: it must reproduce scipy.ndimage.zoom(order=0), the
    function the reference calls (utils.py:360), including the zero-filled last line some size pairs produce."""
    import scipy.ndimage as ndi
    from sln_amodal_b200.targets import zoom_index_map
    rng = np.random.default_rng(0)
    quirks = 0
    for _ in range(200):
        h, w = int(rng.integers(1, 500)), int(rng.integers(1, 500))
        sy = float(rng.choice([1024 / h, rng.uniform(0.3, 3.0)]))
        sx = float(rng.choice([1024 / w, rng.uniform(0.3, 3.0)]))
        oh, ow = int(round(h * sy)), int(round(w * sx))
        if oh == 0 or ow == 0:
            continue
        a = rng.integers(1, 255, (h, w)).astype(np.uint8)
        iy, ix = zoom_index_map(h, oh), zoom_index_map(w, ow)
        got = a[np.maximum(iy, 0)][:, np.maximum(ix, 0)]
        got[iy < 0, :] = 0
        got[:, ix < 0] = 0
        quirks += int((iy < 0).any() or (ix < 0).any())
        assert np.array_equal(got, ndi.zoom(a, zoom=[sy, sx], order=0))
    assert quirks > 0          # the sweep does contain overshooting last lines


def test_after_the_path_entry_points_validate_and_refuse_cpu(built_lib):
    """Mask paste, RPN re-layout and image resize: argument validation before any CUDA call, workspace
    query, no CPU fallback, and install() rebinding MaskRCNN.unmold_detections."""
    import ctypes as C
    import types
    import sln_amodal_b200 as s
    from sln_amodal_b200 import _lib
    L = _lib.lib()
    assert L.sln_unmold_masks(None, 1, 0, 28, None, 64, 64, None, None) == -1 and b"unmold" in L.sln_last_error_string()
    assert L.sln_unmold_masks(None, 1, 300, 300, None, 64, 64, None, None) == -1
    assert L.sln_unmold_masks(None, 0, 28, 28, None, 64, 64, None, None) == 0              # nothing to do
    hs, ws = (C.c_int * 2)(4, 2), (C.c_int * 2)(4, 2)
    assert L.sln_rpn_pack(None, None, hs, ws, 9, 1, 3, 0, None, None, None, None) == -1 and b"levels" in L.sln_last_error_string()
    assert L.sln_rpn_pack(None, None, hs, ws, 2, 1, 9, 0, None, None, None, None) == -1 and b"anchors per location" in L.sln_last_error_string()
    assert L.sln_rpn_pack(None, None, hs, ws, 2, 0, 3, 0, None, None, None, None) == 0    # empty batch
    need = L.sln_resize_image_workspace_bytes(1440, 1920, 3, 1024, 1024)
    assert 1440 * 1024 * 3 <= need <= 1440 * 1024 * 3 + (1 << 20)                          # the 8-bit intermediate + tables
    assert L.sln_resize_image_u8(None, 4, 4, 3, 8, 8, None, None, 0, None) == -1
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SlnError):
            s.unmold_masks(np.zeros((1, 28, 28), np.float32), np.array([[0, 0, 4, 4]]), (8, 8))
        with pytest.raises(_lib.SlnError):
            s.resize_image_device(np.zeros((4, 4, 3), np.uint8), (8, 8))
    with pytest.raises(_lib.SlnError):
        s.rpn_pack([torch.zeros(1, 6, 4, 4)], [torch.zeros(1, 12, 4, 4)])

    model = types.ModuleType("model")

    class MaskRCNN:
        def unmold_detections(self, detections, mrcnn_mask, image_shape, window):
            return "reference"
    model.MaskRCNN = MaskRCNN
    done = s.install(model, None, None)
    assert "model.MaskRCNN.unmold_detections" in done and MaskRCNN.unmold_detections.__name__ == "_unmold_detections"


def test_refine_detections_has_no_cpu_path():
    from sln_amodal_b200 import refine_detections

    class Cfg:
        RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
        IMAGE_SHAPE = np.array([1024, 1024, 3])
        USE_NMS = False

    with pytest.raises(RuntimeError):
        refine_detections(torch.zeros(3, 4), torch.zeros(3, 2), torch.zeros(3, 2, 4), (0, 0, 1, 1), Cfg())


def test_npz_reader_is_bit_identical_to_numpy(tmp_path):
    """npz.read_member (zip member inflated straight into one host buffer) against np.load: stored and deflated archives,
    C and Fortran order, the uint64 label maps (as int64 bit patterns) and an ordinary dtype"""
    from sln_amodal_b200 import npz
    rng = np.random.default_rng(3)
    label = rng.integers(0, 2 ** 63, (37, 53), dtype=np.uint64) | (np.uint64(1) << np.uint64(63))
    other = rng.standard_normal((5, 7, 3)).astype(np.float32)
    for k, save in enumerate((np.savez, np.savez_compressed)):
        p = str(tmp_path / ("a%d.npz" % k))
        save(p, layer=label, other=other, f=np.asfortranarray(label))
        ref = np.load(p)
        got = npz.read_member(p, "layer", pinned=False)
        assert got.dtype == torch.int64 and np.array_equal(got.numpy().view(np.uint64), ref["layer"])
        assert np.array_equal(npz.read_member(p, "other", pinned=False).numpy(), ref["other"])
        assert np.array_equal(npz.read_member(p, "f", pinned=False).numpy().view(np.uint64), ref["f"])
        assert np.array_equal(npz.load_layer_label(p, device="cpu").numpy().view(np.uint64), ref["layer"])
    with pytest.raises(TypeError):
        npz.load_layer_label(p, device="cpu", name="other")
    with pytest.raises(KeyError):
        npz.read_member(p, "missing", pinned=False)

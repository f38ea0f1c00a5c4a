"""The banded EDT's per-column logic (sln_amodal_b200/csrc/edt_band.cuh: the functions the CUDA kernels call) run
lane by lane on the CPU by tests/host_sim/edt_band_sim.cpp, held bit-exact to the oracle.  CPU-only: the simulation is
test infrastructure (it lets the algorithm be checked where there is no GPU), never a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle
from sln_amodal_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("edt_band_sim") / "libedt_band_sim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host_sim", "edt_band_sim.cpp")],
                   check=True)
    lib = C.CDLL(so)

    def run(m):
        m = np.ascontiguousarray(m, np.uint8)
        H, W = m.shape
        out = np.full((H, W), -7, np.int32)
        mt = C.c_int(0)
        rc = lib.edt_band_sim(m.ctypes.data_as(C.c_void_p), H, W, out.ctypes.data_as(C.c_void_p), C.byref(mt))
        assert rc == 0, rc
        assert mt.value <= 34
        return out
    return run


def test_band_sim_shapes(sim):
    rng = np.random.default_rng(0)
    for H, W, p in [(1, 32, 0.5), (5, 64, 0.7), (33, 32, 0.9), (64, 64, 0.999), (100, 128, 0.5), (70, 1024, 0.98),
                    (257, 96, 0.995), (200, 256, 0.9999)]:
        m = (rng.random((H, W)) < p).astype(np.uint8)
        assert np.array_equal(sim(m), oracle.edt_sq(m)), (H, W, p)
    for H, W in [(9, 32), (64, 64), (100, 160)]:
        m = np.ones((H, W), np.uint8)
        assert np.all(sim(m) == (H + W) ** 2)
        m.flat[rng.integers(0, H * W)] = 0
        assert np.array_equal(sim(m), oracle.edt_sq(m))


def test_band_sim_blobs(sim):
    yy, xx = np.mgrid[0:300, 0:320]
    cases = [(((yy - 150) / 140.0) ** 2 + ((xx - 160) / 100.0) ** 2 <= 1).astype(np.uint8)]
    for sl in [(slice(10, 290), slice(40, 300)), (slice(None), slice(40, 300)), (slice(50, 250), slice(None)),
               (slice(31, 33), slice(0, 320)), (slice(32, 64), slice(32, 64)), (slice(0, 300), slice(0, 320))]:
        m = np.zeros((300, 320), np.uint8)
        m[sl] = 1
        cases.append(m)
    m = np.zeros((300, 320), np.uint8)
    m[:, ::37] = 1
    m[::53, :] = 1
    cases.append(m)
    for i, m in enumerate(cases):
        assert np.array_equal(sim(m), oracle.edt_sq(m)), i
        assert np.array_equal(sim(1 - m), oracle.edt_sq(1 - m)), i


def test_band_sim_label_planes(sim):
    """planes of the config-4 generator at 512^2 (every visible plane of two label maps)"""
    for seed in range(2):
        lab = synth.label_map(512, 512, n=6, seed=seed, min_piece=16)
        for i in range(6):
            m = ((lab >> np.uint64(i)) & np.uint64(1)).astype(np.uint8)
            assert np.array_equal(sim(m), oracle.edt_sq(m)), (seed, i)


def test_band_sim_random_unions(sim):
    rng = np.random.default_rng(7)
    for it in range(120):
        H = int(rng.integers(1, 200))
        W = 32 * int(rng.integers(1, 6))
        m = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(1, 5))):
            y0, y1 = sorted(rng.integers(0, H + 1, 2))
            x0, x1 = sorted(rng.integers(0, W + 1, 2))
            if it % 3 == 0:
                m[y0:y1, x0:x1] = 1
            elif it % 3 == 1:
                yy, xx = np.mgrid[0:H, 0:W]
                cy, cx, ry, rx = (y0 + y1) / 2, (x0 + x1) / 2, max(1, (y1 - y0) / 2), max(1, (x1 - x0) / 2)
                m |= ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1).astype(np.uint8)
            else:
                m[y0:y1, x0:x1] = rng.random((y1 - y0, x1 - x0)) < 0.95
        if it % 7 == 0:
            m = 1 - m
        assert np.array_equal(sim(m), oracle.edt_sq(m)), it


def test_band_sim_field_limits_and_long_chains(sim):
    """Row distances up to 1023 (W = 1024, one zero column), a single zero in the far corner (every other band has no
    site at all and only passes the search on: the longest chain of look-ups), rows up to 2047 (H = 2048: the 11-bit
    entry fields at their limit) on a thin map."""
    m = np.ones((256, 1024), np.uint8)
    m[:, 0] = 0
    assert np.array_equal(sim(m), oracle.edt_sq(m))
    m = np.ones((512, 1024), np.uint8)
    m[511, 1023] = 0
    assert np.array_equal(sim(m), oracle.edt_sq(m))
    m = np.ones((2048, 64), np.uint8)
    m[2047, 63] = 0
    assert np.array_equal(sim(m), oracle.edt_sq(m))
    m = np.ones((2048, 32), np.uint8)
    m[0, 0] = 0
    m[1000:1003, 5:9] = 0
    assert np.array_equal(sim(m), oracle.edt_sq(m))

"""CPU tier: the marching-squares tables and contour-[0] walk the CUDA trace kernel instantiates
(csrc/trace_core.h), compiled for the host by tests/host/trace_check.cpp, against the oracle."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import contours_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ unavailable")
    so = str(tmp_path_factory.mktemp("trace") / "_trace_check.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "tests", "host", "trace_check.cpp"),
                    "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.trace_check.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_int)]

    def run(mask):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        h, w = m.shape
        out = np.zeros(4 * h * w + 8, np.uint32)
        closed = ctypes.c_int(0)
        n = lib.trace_check(m.ctypes.data, h, w, out.ctypes.data, out.size, ctypes.byref(closed))
        v = out[:n]
        return np.stack([v >> 16, v & 0xffff], 1).astype(np.int64), bool(closed.value)
    return run


def _same(ref, got, closed):
    assert sorted(map(tuple, ref.tolist())) == sorted(map(tuple, got.tolist()))
    assert closed == (len(ref) > 2 and tuple(ref[0]) == tuple(ref[-1]))


def test_random_masks(harness):
    rng = np.random.default_rng(5)
    for _ in range(1500):
        h, w = rng.integers(2, 14, size=2)
        m = (rng.random((h, w)) < rng.choice([.15, .3, .5, .7, .85])).astype(np.uint8)
        got, closed = harness(m)
        try:
            ref = co.first_contour_lattice(m)
        except IndexError:
            assert len(got) == 0
            continue
        _same(ref, got, closed)


def test_every_2x2_and_3x3_mask_exhaustively(harness):
    for bits in range(1 << 9):
        m = np.array([(bits >> i) & 1 for i in range(9)], np.uint8).reshape(3, 3)
        got, closed = harness(m)
        try:
            ref = co.first_contour_lattice(m)
        except IndexError:
            assert len(got) == 0
            continue
        _same(ref, got, closed)


def test_layer_and_lesion_shapes(harness):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import synth
    yt, _ = synth.layered_pair(1, 96, 128, 6, seed=3, noise=0.01)
    for c in range(6):
        got, closed = harness(yt[0] == c)
        _same(co.first_contour_lattice(yt[0] == c), got, closed)
    lt, _ = synth.lesion_pair(1, 96, 96, 4, seed=4, single_blob_interior=False)
    for c in range(4):
        got, closed = harness(lt[0] == c)
        _same(co.first_contour_lattice(lt[0] == c), got, closed)

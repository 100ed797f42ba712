"""CPU check of the order in which the pair kernel walks the (agent, slot) pairs of a rank: csrc/lqro_enum.cuh compiled with g++
(tests/enum_host/enum_host.cpp).  Whatever the chunk widths and the segment split, every pair must come exactly once, padding
indices must be rejected, a 32-pair batch must stay inside one chunk of one segment with ascending agents (the pair kernel waits
for "rows [0, max agent of the batch]" of a neighbour table that is still being uploaded, lqro_step.cuh), and the first segment
must be finished before the second starts.  The reference's own order is agent after agent (MAIN:1376-1379); results do not
depend on the order (tests/test_gpu_edges.py runs the device side)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "enum_host", "enum_host.cpp")
SO = os.path.join(HERE, "enum_host", "libenumhost.so")
CSRC = os.path.join(HERE, "..", "lqr_obstacles_b200", "csrc")
DEPS = [SRC, os.path.join(HERE, "hull_host", "warp_emul.h"), os.path.join(CSRC, "lqro_enum.cuh"), os.path.join(CSRC, "lqro_common.cuh")]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in DEPS):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-o", SO, SRC])
    lib = ctypes.CDLL(SO)
    lib.enumhost_walk.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 2 + [ctypes.c_longlong]
    lib.enumhost_walk.restype = ctypes.c_longlong
    lib.enumhost_prefix.argtypes = [ctypes.c_int] * 4
    lib.enumhost_prefix.restype = ctypes.c_longlong
    lib.enumhost_batches.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 2
    return lib


def walk(lib, n, K, first, split):
    cap = (n * K + 32 * 2 * 16) + 64
    il = np.full(cap, -2, dtype=np.int32)
    kk = np.full(cap, -2, dtype=np.int32)
    total = lib.enumhost_walk(n, K, first, split, il.ctypes.data_as(ctypes.c_void_p), kk.ctypes.data_as(ctypes.c_void_p), cap)
    assert total <= cap and total % 32 == 0
    return il[:total], kk[:total]


CASES = [(n, K, first, split) for n, K in ((1, 1), (5, 3), (33, 7), (64, 64), (100, 32), (257, 48), (1000, 64), (96, 128))
         for first in (1, 3, 4, 32) for split in (0, 32, 64) if split < n or split == 0]


@pytest.mark.parametrize("n,K,first,split", CASES)
def test_every_pair_once(lib, n, K, first, split):
    il, kk = walk(lib, n, K, first, split)
    real = il >= 0
    assert real.sum() == n * K
    assert np.all((kk >= 0) == real)
    seen = np.zeros((n, K), dtype=np.int32)
    np.add.at(seen, (il[real], kk[real]), 1)
    assert np.all(seen == 1)
    # batches: one segment, ascending agents; the first segment is done before the second starts
    if split > 0:
        in_a = il[real] < split
        first_b = np.argmax(~in_a) if (~in_a).any() else len(in_a)
        assert in_a[:first_b].all() and not in_a[first_b:].any()
    for b in range(0, len(il), 32):
        bi = il[b:b + 32]
        bi = bi[bi >= 0]
        if len(bi) == 0:
            continue
        assert np.all(np.diff(bi) >= 0)
        if split > 0:
            assert (bi < split).all() or (bi >= split).all()
    # padding only at the end of a chunk: a real pair never follows padding inside a batch
    pad = ~real
    for b in np.nonzero(pad)[0]:
        assert pad[b:(b // 32 + 1) * 32].all()


def test_nearest_slots_first(lib):
    """slot-major: every agent's slot s comes before anybody's slot of a later chunk; prefix() counts the padded pairs of the
    chunks that start below a slot"""
    n, K = 200, 64
    il, kk = walk(lib, n, K, 1, 0)
    real = kk >= 0
    chunk_of = np.searchsorted(np.array([1, 2, 4, 8, 16, 32, 64]), kk[real], side="right")
    assert np.all(np.diff(chunk_of) >= 0)
    pad = lambda v: (v + 31) // 32 * 32
    assert lib.enumhost_prefix(n, K, 1, 8) == pad(n) + pad(n) + pad(2 * n) + pad(4 * n)  # chunks 0 | 1 | 2-3 | 4-7


@pytest.mark.parametrize("n,K,first,split", [c for c in CASES if c[0] >= 33])
@pytest.mark.parametrize("width", [32, 8, 2])
def test_batches_tile_the_enumeration(lib, n, K, first, split, width):
    """the batches of a launch (BatchPlan: narrow ones for the chunks below slot 8 of either segment, 32 pairs otherwise) cover
    every index of the enumeration exactly once, in order, never straddle two segments, and every batch number past the end is
    rejected"""
    il, kk = walk(lib, n, K, first, split)
    total = len(il)
    cap = total // 2 + 64
    base = np.zeros(cap, dtype=np.int32)
    bw = np.zeros(cap, dtype=np.int32)
    nb = lib.enumhost_batches(n, K, first, split, 8, width, base.ctypes.data_as(ctypes.c_void_p), bw.ctypes.data_as(ctypes.c_void_p), cap, 5)
    assert 0 < nb < cap
    base, bw = base[:nb], bw[:nb]
    assert base[0] == 0 and np.all(base[1:] == base[:-1] + bw[:-1]) and base[-1] + bw[-1] == total
    assert set(np.unique(bw)) <= {32, width}
    # narrow batches exactly where the chunk starts below slot 8 (padding belongs to its chunk)
    if width < 32:
        pad = lambda v: (v + 31) // 32 * 32
        segs = [(0, n)] if split == 0 else [(0, split), (split, n - split)]
        off = 0
        for _, cnt in segs:
            narrow = lib.enumhost_prefix(cnt, K, first, 8)
            seg_total = lib.enumhost_prefix(cnt, K, first, K)
            in_seg = (base >= off) & (base < off + seg_total)
            assert np.all(bw[in_seg & (base < off + narrow)] == width) and np.all(bw[in_seg & (base >= off + narrow)] == 32)
            off += seg_total
        assert off == total
    # a batch holds ascending agents of one segment
    for b0, w in zip(base, bw):
        bi = il[b0:b0 + w]
        bi = bi[bi >= 0]
        assert np.all(np.diff(bi) >= 0)
        if split > 0 and len(bi):
            assert (bi < split).all() or (bi >= split).all()

"""GPU parity tests, one device stage at a time, through the C ABI's stage-level entry points.
Each compares the CUDA result with the C oracle (oracle/meryl_oracle.c, itself pinned to the
reference binary) on the same seeded input; everything is integer work so the bar is bit-exact."""
import ctypes
import os

import numpy as np
import pytest

import util
import meryl_b200 as mb
from meryl_b200 import synth

pytestmark = pytest.mark.gpu


def _extract_sorted(bases: bytes, k: int, mode: int):
    L = mb.lib()
    n = ctypes.c_uint64(0)
    cap = max(len(bases), 1)
    lo = np.zeros(cap, dtype=np.uint64)
    hi = np.zeros(cap, dtype=np.uint64)
    rc = L.mb_dbg_extract_sorted(k, mode, bases, len(bases), lo.ctypes.data, hi.ctypes.data, cap, ctypes.byref(n))
    assert rc == 0, L.mb_last_error().decode()
    return hi[: n.value], lo[: n.value]


def _oracle_sorted(bases: bytes, k: int, mode: int):
    hi, lo = util.oracle_extract(bases + b".", k, mode)
    order = np.lexsort((lo, hi))
    return hi[order], lo[order]


def _edge_stream(rng, n):
    """Bases with lowercase, N, IUPAC, record breaks, short runs."""
    s = synth.random_bases(n, rng).copy()
    low = rng.random(n) < 0.2
    s[low] |= 0x20
    for ch, frac in ((ord("N"), 0.01), (ord("."), 0.01), (ord("R"), 0.003), (ord("n"), 0.003)):
        s[rng.random(n) < frac] = ch
    return s.tobytes()


@pytest.mark.parametrize("k", [3, 4, 5, 6, 8, 15, 16, 21, 22, 31, 32, 33, 40, 51, 63, 64])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_extract_matches_oracle(k, mode):
    rng = np.random.default_rng(1000 * k + mode)
    bases = _edge_stream(rng, 50_000 if k > 8 else 3_000)
    ghi, glo = _extract_sorted(bases, k, mode)
    ohi, olo = _oracle_sorted(bases, k, mode)
    assert len(glo) == len(olo), (len(glo), len(olo))
    bad = np.nonzero((glo != olo) | (ghi != ohi))[0]
    assert bad.size == 0, "first mismatch at %d: gpu %x:%x oracle %x:%x" % (bad[0], ghi[bad[0]], glo[bad[0]], ohi[bad[0]], olo[bad[0]])


def test_extract_reads_shape_and_tile_boundaries():
    # 150 bp reads separated by '.', long enough to cross many 8192-position tiles and one staging flush boundary
    rng = np.random.default_rng(7)
    reads = synth.random_bases(151 * 40_000, rng).reshape(40_000, 151).copy()
    reads[:, 150] = ord(".")
    bases = reads.tobytes()
    for k in (21, 51):
        ghi, glo = _extract_sorted(bases, k, 0)
        ohi, olo = _oracle_sorted(bases, k, 0)
        assert len(glo) == 40_000 * (150 - k + 1)
        assert np.array_equal(glo, olo) and np.array_equal(ghi, ohi)


def test_extract_empty_and_no_kmers():
    for bases in (b"", b".", b"ACGT", b"N" * 100, b"ACGTACGTAC.ACGTACGTAC"):
        ghi, glo = _extract_sorted(bases, 21, 0)
        assert len(glo) == 0


_SHAPES = [2, 3]


@pytest.fixture(params=_SHAPES, ids=["sweep_128x8", "sweep_256x6"])
def sort_version(request):
    """Both sweep shapes (csrc/sort.cu): 128 threads x 8-bit digits, 256 threads x 6-bit digits."""
    L = mb.lib()
    L.mb_dbg_sort_version.argtypes = [ctypes.c_int]
    old = L.mb_dbg_sort_version(request.param)
    yield request.param
    L.mb_dbg_sort_version(old)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 4095, 4096, 4097, 100_000, 1_000_003])
@pytest.mark.parametrize("bits", [1, 7, 8, 9, 32, 36, 56, 64])
def test_sort_u64(n, bits, sort_version):
    rng = np.random.default_rng(n * 131 + bits)
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    if n > 10:
        keys[: n // 3] = keys[n // 2]          # heavy duplicates
    ref = np.sort(keys)
    got = keys.copy()
    L = mb.lib()
    rc = L.mb_dbg_sort(got.ctypes.data, None, n, bits)
    assert rc == 0, L.mb_last_error().decode()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", [0, 1, 33, 8191, 8192, 8193, 100_000, 3_000_001])
@pytest.mark.parametrize("bits", [1, 8, 9, 22, 32])
def test_sort_u32(n, bits, sort_version):
    rng = np.random.default_rng(n * 7 + bits)
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    if bits < 32:
        keys &= np.uint32((1 << bits) - 1)
    if n > 10:
        keys[: n // 3] = keys[n // 2]
    ref = np.sort(keys)
    got = keys.copy()
    L = mb.lib()
    rc = L.mb_dbg_sort32(got.ctypes.data, n, bits)
    assert rc == 0, L.mb_last_error().decode()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 300_001])
@pytest.mark.parametrize("bits", [65, 70, 92, 118, 128])
def test_sort_u128(n, bits, sort_version):
    rng = np.random.default_rng(n * 17 + bits)
    lo = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    hi = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if bits < 128:
        hi &= np.uint64((1 << (bits - 64)) - 1)
    if n > 10:
        hi[: n // 2] = hi[0]
        lo[: n // 4] = lo[0]
    order = np.lexsort((lo, hi))
    glo, ghi = lo.copy(), hi.copy()
    L = mb.lib()
    rc = L.mb_dbg_sort(glo.ctypes.data, ghi.ctypes.data, n, bits)
    assert rc == 0, L.mb_last_error().decode()
    assert np.array_equal(glo, lo[order]) and np.array_equal(ghi, hi[order])


def _encode_gpu(k, wp, prefix, lo, hi, vals):
    L = mb.lib()
    n = len(lo)
    cap = 4096 + 64 * n
    out = np.zeros(cap, dtype=np.uint8)
    olen = ctypes.c_uint64(0)
    lo = np.ascontiguousarray(lo, dtype=np.uint64)
    hi = np.ascontiguousarray(hi, dtype=np.uint64)
    vals = np.ascontiguousarray(vals, dtype=np.uint32)
    rc = L.mb_dbg_encode_block(k, wp, prefix, n, lo.ctypes.data, hi.ctypes.data, vals.ctypes.data,
                               out.ctypes.data, cap, ctypes.byref(olen))
    assert rc == 0, L.mb_last_error().decode()
    return out[: olen.value].tobytes()


def _random_block(rng, w_suffix, n):
    """n distinct sorted suffixes of w_suffix bits as (hi, lo) uint64 arrays."""
    space = 1 << w_suffix
    n = min(n, space)
    if w_suffix <= 62:
        vals = rng.choice(space, size=n, replace=False) if space <= 4 * max(n, 1) else np.unique(rng.integers(0, space, size=int(n * 1.2) + 8))[:n]
        vals = np.sort(np.asarray(vals, dtype=np.uint64))
        n = len(vals)
        return np.zeros(n, np.uint64), vals
    hi = rng.integers(0, 1 << (w_suffix - 64), size=n, dtype=np.uint64) if w_suffix > 64 else np.zeros(n, np.uint64)
    lo = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if w_suffix < 64:
        lo &= np.uint64((1 << w_suffix) - 1)
    order = np.lexsort((lo, hi))
    hi, lo = hi[order], lo[order]
    keep = np.ones(n, bool)
    keep[1:] = (hi[1:] != hi[:-1]) | (lo[1:] != lo[:-1])
    return hi[keep], lo[keep]


@pytest.mark.parametrize("k,wp", [(21, 10), (21, 24), (31, 14), (32, 10), (33, 10), (51, 10), (51, 23), (64, 20), (8, 6), (16, 12), (5, 6)])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 8, 9, 255, 256, 257, 1000, 20_000])
def test_encode_block_matches_oracle(k, wp, n):
    w_suffix = 2 * k - wp
    rng = np.random.default_rng(k * 1000 + wp * 37 + n)
    hi, lo = _random_block(rng, w_suffix, n)
    vals = rng.integers(1, 50, size=len(lo), dtype=np.uint32)
    if len(vals) > 3:
        vals[1] = 0xFFFFFFFF
        vals[2] = 70_000
    prefix = int(rng.integers(0, 1 << wp))
    want = util.oracle_encode_block(w_suffix, prefix, hi, lo, vals)
    got = _encode_gpu(k, wp, prefix, lo, hi, vals)
    assert len(got) == len(want), (len(got), len(want))
    if got != want:
        a, b = np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8)
        d = np.nonzero(a != b)[0]
        raise AssertionError("block differs at byte %d of %d (%d bytes differ)" % (d[0], len(want), d.size))


def test_encode_dense_block_zero_binary_bits():
    # every suffix present: u == wSuffix, b == 0 (the simple-mode db shape for small k)
    for k, wp in [(8, 6), (5, 6), (7, 6)]:
        w_suffix = 2 * k - wp
        lo = np.arange(1 << w_suffix, dtype=np.uint64)
        hi = np.zeros_like(lo)
        vals = (np.arange(len(lo)) % 7 + 1).astype(np.uint32)
        want = util.oracle_encode_block(w_suffix, 3, hi, lo, vals)
        got = _encode_gpu(k, wp, 3, lo, hi, vals)
        assert got == want


def test_encode_clustered_suffixes_long_unary_runs():
    # suffixes bunched at the two ends of the range: long unary codes, cuts inside the unary / binary / value regions
    rng = np.random.default_rng(99)
    for k, wp, n in [(21, 10, 5000), (31, 10, 3000), (51, 10, 3000)]:
        w_suffix = 2 * k - wp
        hi, lo = _random_block(rng, w_suffix, n)
        if w_suffix <= 64:
            lo = np.sort(np.concatenate([lo[: n // 2] >> np.uint64(12), lo[n // 2:] | np.uint64(((1 << 10) - 1) << (w_suffix - 10))]))
            keep = np.ones(len(lo), bool)
            keep[1:] = lo[1:] != lo[:-1]
            lo, hi = lo[keep], np.zeros(int(keep.sum()), np.uint64)
        vals = rng.integers(1, 1000, size=len(lo), dtype=np.uint32)
        want = util.oracle_encode_block(w_suffix, 1, hi, lo, vals)
        got = _encode_gpu(k, wp, 1, lo, hi, vals)
        assert got == want


def test_count_lookback_warp():
    """The count kernels' warp-wide look-back on prepared status arrays (chains longer than one 32-word window)."""
    from test_cusim_stages import check_lookback
    check_lookback(mb.lib())

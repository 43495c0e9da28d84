"""Single device stages run in the cusim host emulator (see tests/test_cusim_pipeline.py): both sweep
kernels of the segmented radix sort on adversarial key sets, through the same debug entry points the GPU
stage tests use (meryl_b200/csrc/debug.cu)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import util
import meryl_b200 as mb

SIM_DIR = os.path.join(util.ROOT, "tests", "cusim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libmeryl_b200_sim.so")


@pytest.fixture(scope="module", autouse=True)
def sim_library():
    subprocess.run(["make", "-s", "-j8", "-C", SIM_DIR], check=True)
    saved = (mb.LIB_PATH, mb._lib)
    mb.LIB_PATH, mb._lib = SIM_LIB, None
    yield
    mb.LIB_PATH, mb._lib = saved


@pytest.fixture(params=[2, 3], ids=["sweep_128x8", "sweep_256x6"])
def sort_version(request):
    L = mb.lib()
    L.mb_dbg_sort_version.argtypes = [ctypes.c_int]
    old = L.mb_dbg_sort_version(request.param)
    yield request.param
    L.mb_dbg_sort_version(old)


def _keys32(n, bits, kind, seed):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    if kind == "dups" and n > 10:
        keys[: n // 3] = keys[n // 2]
    elif kind == "equal":
        keys[:] = 0xdeadbeef                     # one digit gets the whole tile in every pass
    elif kind == "sorted":
        keys = np.sort(keys)
    elif kind == "reversed":
        keys = np.sort(keys)[::-1].copy()
    if bits < 32:
        keys &= np.uint32((1 << bits) - 1)
    return keys


@pytest.mark.parametrize("n,bits,kind", [(0, 32, "random"), (1, 32, "random"), (33, 9, "random"), (8191, 32, "dups"), (8192, 32, "equal"),
                                         (8193, 22, "random"), (30_000, 32, "dups"), (20_000, 32, "reversed"), (17_000, 8, "sorted")])
def test_sim_sort_u32(n, bits, kind, sort_version):
    keys = _keys32(n, bits, kind, n * 7 + bits)
    got = keys.copy()
    L = mb.lib()
    rc = L.mb_dbg_sort32(got.ctypes.data, n, bits)
    assert rc == 0, L.mb_last_error().decode()
    assert np.array_equal(got, np.sort(keys))


@pytest.mark.parametrize("n,bits", [(2, 64), (4095, 36), (4096, 56), (4097, 64), (13_000, 52), (9_000, 7)])
def test_sim_sort_u64(n, bits, sort_version):
    rng = np.random.default_rng(n * 131 + bits)
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    if n > 10:
        keys[: n // 3] = keys[n // 2]
    got = keys.copy()
    L = mb.lib()
    rc = L.mb_dbg_sort(got.ctypes.data, None, n, bits)
    assert rc == 0, L.mb_last_error().decode()
    assert np.array_equal(got, np.sort(keys))


@pytest.mark.parametrize("n,bits", [(1, 128), (2047, 65), (2048, 92), (2049, 128), (7_000, 118)])
def test_sim_sort_u128(n, bits, sort_version):
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


def _lookback_cases():
    rng = np.random.default_rng(3)
    cases = []
    for n, inc_every in [(1, 1), (5, 2), (31, 40), (32, 40), (33, 40), (64, 70), (65, 33), (200, 64), (1000, 1001), (1000, 7), (97, 1)]:
        vals = rng.integers(0, 1 << 20, size=n, dtype=np.uint64)
        flags = np.ones(n, dtype=np.uint64)                       # aggregates ...
        flags[::inc_every] = 2                                   # ... with an inclusive prefix now and then (tile 0 always)
        cases.append((vals, flags))
    return cases


def check_lookback(L):
    """Expected: walking down from tile-1, add aggregates until (and including) the first inclusive prefix."""
    L.mb_dbg_lookback.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    for vals, flags in _lookback_cases():
        status = (flags << np.uint64(62)) | vals
        n = len(vals)
        for tile in sorted({0, 1, n // 2, n - 1, n}):
            want, j = 0, tile - 1
            while j >= 0:
                want += int(vals[j])
                if flags[j] == 2:
                    break
                j -= 1
            got = ctypes.c_uint64(0)
            rc = L.mb_dbg_lookback(status.ctypes.data, n, tile, ctypes.byref(got))
            assert rc == 0, L.mb_last_error().decode()
            assert got.value == want, (n, tile, got.value, want)


def test_sim_lookback_warp():
    check_lookback(mb.lib())

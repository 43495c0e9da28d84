"""Debug helper (GPU box): per-phase clock breakdown of the sort sweep and extract kernels."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meryl_b200 as mb

L = mb.lib()
L.mb_dbg_sort_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.mb_dbg_extract_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]


def report(title, out, names):
    tiles = max(1, out[0])
    tot = sum(out[1:1 + len(names)])
    print(title, "tiles", out[0], "clk/tile", round(tot / tiles))
    for i, nme in enumerate(names):
        print("   %-28s %9.0f clk/tile %5.1f%%" % (nme, out[1 + i] / tiles, 100.0 * out[1 + i] / max(1, tot)))


def sort_phases(n):
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64)
    out = (ctypes.c_uint64 * 8)()
    L.mb_dbg_sort_profile(None, 1)
    assert L.mb_dbg_sort(keys.ctypes.data, None, n, 32) == 0
    L.mb_dbg_sort_profile(out, 0)
    report("sort_sweep", list(out), ["tile id", "load+rank", "scan+reorder", "own lookback", "cta lookback wait", "copy-out"])
    assert np.all(keys[1:] >= keys[:-1])


def sort32_phases(n):
    L.mb_dbg_sort32.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32]
    rng = np.random.default_rng(2)
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    out = (ctypes.c_uint64 * 8)()
    L.mb_dbg_sort_profile(None, 1)
    assert L.mb_dbg_sort32(keys.ctypes.data, n, 32) == 0
    L.mb_dbg_sort_profile(out, 0)
    report("sort_sweep<u32>", list(out), ["tile id", "load+rank", "scan+reorder", "own lookback", "cta lookback wait", "copy-out"])
    assert np.all(keys[1:] >= keys[:-1])


def extract_phases(gbp):
    import torch
    n_reads = int(gbp * 1e9) // 150
    rec0 = mb.synth_record_bytes(150, 0)
    dev = torch.empty(n_reads * rec0, dtype=torch.uint8, device="cuda")
    mb.synth_reads_device(dev.data_ptr(), 0, n_reads, 150, 3_000_000_000, 0.001, 2, 0)
    torch.cuda.synchronize()
    out = (ctypes.c_uint64 * 8)()
    L.mb_dbg_extract_profile(None, 1)
    with mb.Counter(21, 18, flags=mb.FLAG_NO_IMAGES) as c:
        c.add_bases_device(dev.data_ptr(), n_reads * rec0)
        c.count()
        st = c.stats()
    L.mb_dbg_extract_profile(out, 0)
    report("extract<scatter>", list(out), ["wait TMA tile", "own phase A", "cta phase A", "own reserve", "cta reserve", "own phase B", "cta phase B"])
    print("   stage ms:", {k: round(getattr(st, k), 2) for k in ("msHistogram", "msExtract", "msSortHist", "msSort", "msCount", "msEncode")})


if __name__ == "__main__":
    sort_phases(int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000)
    sort32_phases(int(float(sys.argv[1])) * 2 if len(sys.argv) > 1 else 400_000_000)
    extract_phases(float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)

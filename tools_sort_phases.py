"""Debug helper: per-phase clock breakdown of the sort sweep kernel (run on a GPU box)."""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import meryl_b200 as mb
L = mb.lib()
L.mb_dbg_sort_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
rng = np.random.default_rng(1)
keys = rng.integers(0, 2**32, size=n, dtype=np.uint64)
out = (ctypes.c_uint64 * 8)()
L.mb_dbg_sort_profile(None, 1)
assert L.mb_dbg_sort(keys.ctypes.data, None, n, 32) == 0
L.mb_dbg_sort_profile(out, 1)
tiles = out[0]
names = ["tiles", "tile id", "load+rank", "scan+reorder", "own lookback", "cta lookback wait", "copy-out"]
tot = sum(out[1:7])
for i in range(1, 7):
    print("%-20s %10.0f clk/tile  %5.1f%%" % (names[i], out[i] / tiles, 100 * out[i] / tot))
print("tiles", tiles, "total clk/tile", tot / tiles)
assert np.all(keys[1:] >= keys[:-1])

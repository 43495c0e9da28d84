#!/usr/bin/env python
"""A/B of the device-resident count in ONE process: sweep kernel v1/v2 x extraction filter on/off, on the bench
workload (k=21, 150 bp synthetic reads) at a reduced size with the number of passes forced like the 100 Gbp run.
usage: tools/ab_bench.py [gbp] [passes]      -> one JSON line per configuration (stage times in ms)"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import meryl_b200 as mb  # noqa: E402

K, READ_LEN, GENOME_LEN, ERR, SEED = 21, 150, 3_000_000_000, 0.001, 2
gbp = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n_reads = int(gbp * 1e9) // READ_LEN
n_kmers = n_reads * (READ_LEN - K + 1)
rec0 = mb.synth_record_bytes(READ_LEN, 0)
pl = mb.plan(K, int(100e9) // READ_LEN * mb.synth_record_bytes(READ_LEN, 1), 1024 << 30)   # the 100 Gbp run's prefix width
chunk = 4_000_000
dev = torch.empty(chunk * rec0, dtype=torch.uint8, device="cuda")
L = mb.lib()
L.mb_dbg_sort_version.argtypes = [ctypes.c_int]
if passes > 1:
    os.environ["MB_MAX_PASS_KEYS"] = str(int(n_kmers / passes * 1.02))

configs = [(3, 1), (3, 0), (2, 1), (2, 0)]
if len(sys.argv) > 3 and sys.argv[3] == "sortonly":          # the sweep kernel shapes, filter on
    configs = [(int(v), 0) for v in os.environ.get("MB_AB_SHAPES", "2,3").split(",")]
extra = [{}]
if len(sys.argv) > 3 and sys.argv[3] == "env":               # any list of environments: MB_AB_ENVS='[{}, {"MB_X": "1"}, ...]' (default sweep)
    configs = [(L.mb_dbg_sort_version(0), 0)]
    extra = json.loads(os.environ.get("MB_AB_ENVS", "[{}]"))
touched = set(k for e in extra for k in e)
for (sort_v, nofilter), env in [(c, e) for c in configs for e in extra]:
    for k in touched:
        os.environ.pop(k, None)
    os.environ.update(env)
    L.mb_dbg_sort_version(sort_v)
    if nofilter:
        os.environ["MB_NO_PREFILTER"] = "1"
    else:
        os.environ.pop("MB_NO_PREFILTER", None)
    best = None
    for it in range(2):
        c = mb.Counter(K, pl.w_prefix, "count", flags=mb.FLAG_NO_IMAGES)
        try:
            for first in range(0, n_reads, chunk):
                m = min(chunk, n_reads - first)
                mb.synth_reads_device(dev.data_ptr(), first, m, READ_LEN, GENOME_LEN, ERR, SEED, 0)
                torch.cuda.synchronize()
                c.add_bases_device(dev.data_ptr(), m * rec0)
            torch.cuda.synchronize()
            c.count()
            st = c.stats()
            assert st.nKmers == n_kmers
            if best is None or st.msTotal < best.msTotal:
                best = st
        finally:
            c.close()
    d = best.as_dict()
    sort_bytes = d["sortLaunchKeys"] * 2 * d["keyBytes"]
    print(json.dumps({"sort": sort_v, "prefilter": not nofilter, "env": env, "gbp": gbp, "kmers": n_kmers, "passes": d["nPasses"], "sort_passes": d["sortPasses"],
                      "ms": {k: round(d[k], 2) for k in ("msTotal", "msHistogram", "msExtract", "msSortHist", "msSort", "msCount", "msEncode")},
                      "gkmers_per_s": round(n_kmers / d["msTotal"] / 1e6, 2),
                      "sweep_GBps": round(sort_bytes / (d["msSort"] / 1000) / 1e9, 1) if d["msSort"] else None,
                      "nDistinct": d["nDistinct"], "nTotal": d["nTotal"]}), flush=True)

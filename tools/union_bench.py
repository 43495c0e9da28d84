#!/usr/bin/env python
"""`meryl union-sum` of two count dbs on one B200 next to the reference binary on the host cores (BASELINE.json
configs[4] at a reduced, stated size).  usage: tools/union_bench.py [gbp per input]   -> one JSON line"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import meryl_b200 as mb  # noqa: E402
import util  # noqa: E402

K, READ_LEN, GENOME_LEN, ERR = 21, 150, 3_000_000_000, 0.001
gbp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
n_reads = int(gbp * 1e9) // READ_LEN
rec0 = mb.synth_record_bytes(READ_LEN, 0)
pl = mb.plan(K, n_reads * mb.synth_record_bytes(READ_LEN, 1), 64 << 30)
tmp = tempfile.mkdtemp(prefix="merylunion_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    dbs, distinct = [], []
    dev = torch.empty(n_reads * rec0, dtype=torch.uint8, device="cuda")
    for i, seed in enumerate((11, 12)):
        with mb.Counter(K, pl.w_prefix) as c:
            mb.synth_reads_device(dev.data_ptr(), 0, n_reads, READ_LEN, GENOME_LEN, ERR, seed, 0)
            torch.cuda.synchronize()
            c.add_bases_device(dev.data_ptr(), n_reads * rec0)
            c.count()
            d = os.path.join(tmp, "in%d.meryl" % i)
            c.write_db(d)
            dbs.append(d)
            distinct.append(c.stats().nDistinct)
    del dev
    out = os.path.join(tmp, "gpu.meryl")
    if gbp <= 1.0:                                   # (large runs: no room for a third output in /dev/shm; the CUDA context is warm from the counts)
        mb.union_sum(dbs, os.path.join(tmp, "warm.meryl"))
        shutil.rmtree(os.path.join(tmp, "warm.meryl"), ignore_errors=True)
    t0 = time.perf_counter()
    st = mb.union_sum(dbs, out)
    wall = time.perf_counter() - t0
    res = {"workload": "union-sum of two k=21 count dbs (%d + %d distinct k-mers, prefix width %d) in /dev/shm" % (distinct[0], distinct[1], pl.w_prefix),
           "kmers_in": st.nInputKmers, "kmers_out": st.nOutputKmers, "gpu_wall_s": round(wall, 3), "gpu_device_ms": round(st.msTotal, 1),
           "gpu_stage_ms": {"read+decode": round(st.msDecode, 1), "merge": round(st.msMerge, 1), "encode+write": round(st.msEncode, 1)},
           "gpu_kmers_in_per_s": round(st.nInputKmers / wall), "groups": st.nGroups, "launches": st.nLaunches}
    if util.have_ref():
        ref = os.path.join(tmp, "ref.meryl")
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        util.run_ref(["threads=%d" % cores, "union-sum", *dbs, "output", ref], timeout=900)
        rt = time.perf_counter() - t0
        in_bytes = sum(os.path.getsize(os.path.join(d, f)) for d in dbs for f in os.listdir(d))
        out_bytes = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
        res.update({"db_bytes_in": in_bytes, "db_bytes_out": out_bytes})
        util.assert_same_db(ref, out)
        res.update({"reference_wall_s": round(rt, 3), "reference_cores": cores, "reference_kmers_in_per_s": round(st.nInputKmers / rt),
                    "speedup": round(rt / wall, 1), "parity": "129 files byte-identical to the reference binary's"})
    print(json.dumps(res), flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)

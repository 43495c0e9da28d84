"""Timing of the entry points written without a GPU at hand, next to their CPU counterparts (run on the GPU box):
  a. mb_existence / meryl-lookup-b200 -existence   vs the reference's meryl-lookup (oracle/_ref/meryl-lookup)
  b. mb_add_file: the multi-threaded host parser's rate (round 2 measured a device form of the state machine next to it:
     6.6 GB/s against 14.3 GB/s on 16 host threads -- slower, and removed)
usage: python tools/lookup_bench.py [Mbp of reads, default 300]      -> one JSON line per measurement"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import meryl_b200 as mb                     # noqa: E402
from meryl_b200 import synth                # noqa: E402
import util                                 # noqa: E402

if os.environ.get("MB_BENCH_SIM"):          # dry run of this script on a machine without a GPU (tests/cusim)
    mb.LIB_PATH, mb._lib = os.path.join(ROOT, "tests", "cusim", "_build", "libmeryl_b200_sim.so"), None


def main():
    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
    n_reads = int(mbp * 1e6) // 150
    tmp = tempfile.mkdtemp(prefix="lookupbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fq = os.path.join(tmp, "reads.fq")
    with open(fq, "wb") as fh:
        for part in range(0, n_reads, 200_000):
            fh.write(synth.reads_fastq_bytes(3_000_000, min(200_000, n_reads - part), 150, 0.005, 11 + part, genome_seed=5))
    size = os.path.getsize(fq)
    db = os.path.join(tmp, "reads.meryl")
    # b. ingest + count through the host parser and through the device parser
    for tag, env in (("host_parser", {}),):
        for k, v in env.items():
            os.environ[k] = v
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            with mb.Counter(21, 12, "count") as c:
                c.add_file(fq)
                t_ingest = time.perf_counter() - t0
                c.count()
                if tag == "host_parser":
                    c.write_db(db)
            best = t_ingest if best is None else min(best, t_ingest)
        for k in env:
            del os.environ[k]
        print(json.dumps({"what": "mb_add_file " + tag, "file_bytes": size, "seconds": best, "GB_per_s": size / best / 1e9}), flush=True)
    # a. existence of the reads' k-mers in their own database
    q = os.path.join(tmp, "q.fq")
    with open(q, "wb") as fh:
        fh.write(synth.reads_fastq_bytes(3_000_000, min(n_reads, 400_000), 150, 0.02, 99, genome_seed=5))
    for exe, tag in ((mb.LOOKUP_CLI_PATH, "meryl-lookup-b200"), (util.REF_LOOKUP_BIN, "reference meryl-lookup")):
        if not os.access(exe, os.X_OK):
            continue
        t0 = time.perf_counter()
        env = dict(os.environ, LD_PRELOAD=mb.LIB_PATH) if os.environ.get("MB_BENCH_SIM") and exe == mb.LOOKUP_CLI_PATH else None
        p = subprocess.run([exe, "-threads", "16", "-existence", "-sequence", q, "-mers", db, "-output", os.path.join(tmp, tag.split()[0] + ".txt")], capture_output=True, text=True, env=env)
        dt = time.perf_counter() - t0
        print(json.dumps({"what": tag + " -existence (whole command)", "rc": p.returncode, "query_bases": 150 * min(n_reads, 400_000), "seconds": dt}), flush=True)
    a = open(os.path.join(tmp, "meryl-lookup-b200.txt"), "rb").read() if os.path.exists(os.path.join(tmp, "meryl-lookup-b200.txt")) else b"a"
    b = open(os.path.join(tmp, "reference.txt"), "rb").read() if os.path.exists(os.path.join(tmp, "reference.txt")) else b"b"
    print(json.dumps({"what": "existence outputs identical", "value": a == b}), flush=True)
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()

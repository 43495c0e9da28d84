"""Byte parity AT THE HEADLINE PLAN (bench.py's C2 configuration, BASELINE.json configs[1]): the bench's own synthetic
reads, the prefix width the 100 Gbp job is planned with (wPrefix = 24: 16.8 M blocks), u32 suffix keys, five
file-aligned passes -- the database must be the unmodified reference binary's, byte for byte.  The reads are written by
oracle/synth_reads (the CPU twin of the device generator, checked here too); the reference needs ~125 GB of host memory
at this prefix width (3232 B of bookkeeping + one 4 KB segment per prefix), so the test skips on a small host."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import util
import meryl_b200 as mb

pytestmark = pytest.mark.gpu

SYNTH = os.path.join(util.ORACLE_DIR, "synth_reads")
K, READ_LEN, GENOME_LEN, ERR, SEED = 21, 150, 3_000_000_000, 0.001, 2          # bench.py's constants


def _cpu_reads(path, first, n, fmt, read_len=READ_LEN, err=ERR):
    subprocess.run([SYNTH, path, str(first), str(n), str(read_len), str(GENOME_LEN), repr(err), str(SEED), str(fmt)], check=True)


@pytest.mark.parametrize("fmt", [0, 1])
def test_cpu_generator_equals_device_generator(tmp_path, fmt):
    """oracle/synth_reads.c writes the bytes csrc/synth.cu writes (so the reference arm times the SAME reads)."""
    import torch
    first, n = 123_456_789, 20_011
    rec = mb.synth_record_bytes(READ_LEN, fmt)
    dev = torch.empty((n * rec + 15) // 16 * 16, dtype=torch.uint8, device="cuda")
    mb.synth_reads_device(dev.data_ptr(), first, n, READ_LEN, GENOME_LEN, ERR, SEED, fmt)
    torch.cuda.synchronize()
    p = str(tmp_path / "cpu.txt")
    _cpu_reads(p, first, n, fmt)
    assert np.array_equal(np.fromfile(p, dtype=np.uint8), dev[: n * rec].cpu().numpy())


@pytest.mark.timeout(1500, method="thread")
def test_headline_plan_vs_reference_binary(monkeypatch):
    import psutil
    if not util.have_ref():
        pytest.skip("oracle/_ref/meryl did not travel to this box")
    if psutil.virtual_memory().available < (160 << 30):
        pytest.skip("the reference needs ~125 GB of host memory at wPrefix = 24")
    gbp = float(os.environ.get("MB_HEADLINE_GBP", "1.0"))
    n_reads = int(gbp * 1e9) // READ_LEN
    n_kmers = n_reads * (READ_LEN - K + 1)
    n_est = int(100e9) // READ_LEN * mb.synth_record_bytes(READ_LEN, 1)          # what bench.py plans the 100 Gbp job with
    pl = mb.plan(K, n_est, 1024 << 30)
    assert pl.w_prefix == 24
    tmp = tempfile.mkdtemp(prefix="mbheadline_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fq = os.path.join(tmp, "reads.fq")
        _cpu_reads(fq, 0, n_reads, 1)
        ref, out = os.path.join(tmp, "ref.meryl"), os.path.join(tmp, "gpu.meryl")
        log = util.run_ref(["k=%d" % K, "threads=%d" % min(64, os.cpu_count() or 4), "memory=1024", "n=%d" % n_est, "count", fq, "output", ref], timeout=1200)
        assert "prefix" in log.lower()
        monkeypatch.setenv("MB_MAX_PASS_KEYS", str(int(n_kmers / 5 * 1.02)))     # five passes, like the 100 Gbp job on one B200
        with mb.Counter(K, pl.w_prefix, "count") as c:
            c.add_file(fq)
            c.count()
            st = c.stats()
            c.write_db(out)
        assert st.nKmers == n_kmers
        assert st.keyBytes == 4, "the headline run sorts u32 suffix keys"
        assert st.nPasses >= 5, st.nPasses                                       # (file-aligned cuts: five or six)
        util.assert_same_db(ref, out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

"""GPU parity of the count batches (W6: merylBlockWriter::finishBatch / mergeBatches,
src/utility/src/kmers-v1/kmers-writer-block.C:172-386): reads that outgrow the device are counted in batches cut at record
breaks and the batches are merged on the device (decode -> merge with saturating sums -> encode with the block writer's
rules).  MB_MAX_BATCH_BASES makes small inputs take that road; the db must be the oracle's single-batch db, byte for byte.
Emulator twin: tests/test_cusim_pipeline.py::test_sim_count_batches.  (The file sorts last on purpose: this road was
written after the round's last GPU call.)"""
import os
import subprocess

import numpy as np
import pytest

import util
import meryl_b200 as mb
from meryl_b200 import synth

pytestmark = pytest.mark.gpu


def _count_vs_oracle(tmp_path, files, k, w_prefix, op="count"):
    paths = []
    for name, data in files.items():
        p = tmp_path / name
        p.write_bytes(data)
        paths.append(str(p))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "gpu.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=%d" % w_prefix, op, *paths, "output", ora])
    with mb.Counter(k, w_prefix, op) as c:
        for p in paths:
            c.add_file(p)
        c.count()
        st = c.stats()
        c.write_db(out)
    util.assert_same_db(ora, out)
    return st


@pytest.mark.parametrize("k,w_prefix,op,batch_bases", [(21, 14, "count", 2_000_000), (31, 12, "count-forward", 1_500_000),
                                                       (51, 10, "count", 3_000_000), (9, 6, "count-reverse", 1_000_000)])
def test_count_batches_db_bit_exact(tmp_path, monkeypatch, k, w_prefix, op, batch_bases):
    monkeypatch.setenv("MB_MAX_BATCH_BASES", str(batch_bases))
    reads = synth.reads_fastq_bytes(300_000, 50_000, 150, 0.004, 7 + k)
    st = _count_vs_oracle(tmp_path, {"r.fq": reads, "g.fa": synth.fasta_bytes(3, 400_000, k)}, k, w_prefix, op=op)
    assert st.nBatches >= 3
    assert st.nKmers == 50_000 * (150 - k + 1) + 3 * (400_000 - k + 1) and st.nTotal == st.nKmers


def test_count_batches_pinned_push_and_output_buffer(tmp_path, monkeypatch):
    """One long push from page-locked memory (the 64 MB direct-DMA road stops where a batch is due), output into a
    caller-owned pinned buffer."""
    import torch
    monkeypatch.setenv("MB_MAX_BATCH_BASES", str(120 << 20))
    n_reads, rl, k = 2_400_000, 150, 21
    rec = mb.synth_record_bytes(rl, 0)
    dev = torch.empty(n_reads * rec, dtype=torch.uint8, device="cuda")
    mb.synth_reads_device(dev.data_ptr(), 0, n_reads, rl, 50_000_000, 0.002, 5, 0)
    torch.cuda.synchronize()
    host = torch.empty(n_reads * rec, dtype=torch.uint8, pin_memory=True)
    host.copy_(dev)
    out_buf = torch.empty(3 << 30, dtype=torch.uint8, pin_memory=True)
    pl = mb.plan(k, n_reads * rec, 8 << 30)

    def run(limit):
        if limit:
            monkeypatch.setenv("MB_MAX_BATCH_BASES", str(limit))
        else:
            monkeypatch.delenv("MB_MAX_BATCH_BASES")
        with mb.Counter(k, pl.w_prefix) as c:
            c.set_output_buffer(out_buf.data_ptr(), out_buf.numel())
            c.add_bases_ptr(host.data_ptr(), n_reads * rec, end_of_sequence=True)
            c.count()
            st = c.stats()
            images = [bytes(c.image(i)) for i in range(129)]
        return st, images

    st1, one = run(0)                                          # one batch: the road every other test checks against the oracle
    st3, many = run(120 << 20)
    assert st1.nBatches == 1 and st3.nBatches >= 3
    assert st3.nKmers == st1.nKmers == n_reads * (rl - k + 1)
    assert (st3.nDistinct, st3.nUnique, st3.nTotal) == (st1.nDistinct, st1.nUnique, st1.nTotal)
    assert many == one


def test_cli_count_batches(tmp_path, monkeypatch):
    fq = tmp_path / "r.fq"
    fq.write_bytes(synth.reads_fastq_bytes(100_000, 30_000, 150, 0.003, 21))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "cli.meryl")
    util.run_oracle(["k=21", "memory=2", "count", str(fq), "output", ora])
    env = dict(os.environ, MB_MAX_BATCH_BASES="1000000")
    p = subprocess.run([mb.CLI_PATH, "k=21", "memory=2", "count", str(fq), "output", out], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    util.assert_same_db(ora, out)

"""GPU tests added after round 2's last GPU call: green in the emulator (tests/cusim, also through `MB_TEST_LIB=... pytest -m
gpu`) and in the emulator soak, not yet run on hardware.  The file sorts last so that `pytest -x` reaches every test that has
passed on a B200 before it reaches these."""
import os

import pytest

import util
import meryl_b200 as mb
from meryl_b200 import synth
import union_cases as uc

pytestmark = pytest.mark.gpu


def _count_vs_oracle(tmp_path, files, k, words, counter=None):
    paths = []
    for name, data in files.items():
        p = tmp_path / name
        p.write_bytes(data)
        paths.append(str(p))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "gpu.meryl")
    util.run_oracle(["k=%d" % k] + words + ["count", *paths, "output", ora])
    if counter is None:
        st = mb.count(paths, out, k, "count", memory_gb=2)
    else:
        with mb.Counter(k, counter, "count") as c:
            for p in paths:
                c.add_file(p)
            c.count()
            st = c.stats()
            c.write_db(out)
    util.assert_same_db(ora, out)
    return st


def test_k3_simple_mode_geometry(tmp_path):
    """k = 3: the simple counter's geometry, prefix 6 = the whole k-mer, no suffix bits (merylOp-countSimple.C:196-199)."""
    _count_vs_oracle(tmp_path, {"a.fa": synth.fasta_bytes(2, 30_000, 103)}, 3, ["memory=2"])


def test_k3_count_batches(tmp_path, monkeypatch):
    """... and the merge of its count batches (the merge keeps the count's geometry)."""
    monkeypatch.setenv("MB_MAX_BATCH_BASES", "1000000")
    st = _count_vs_oracle(tmp_path, {"r.fq": synth.reads_fastq_bytes(300_000, 50_000, 150, 0.004, 10), "g.fa": synth.fasta_bytes(3, 400_000, 3)},
                          3, ["wprefix=6"], counter=6)
    assert st.nBatches >= 3


def test_union_k6_no_suffix_bits(tmp_path):
    """Set operations at k = 6: the stream writer's 12-bit prefix is the whole k-mer (the reference binary writes such
    databases; for k = 5 it fails an assertion)."""
    uc.case_k_sweep(str(tmp_path), 6, n=60_000)


@pytest.mark.parametrize("k,limit", [(21, 30), (40, 33), (32, 50)])
def test_stream_block_split(tmp_path, monkeypatch, k, limit):
    uc.case_stream_block_split(str(tmp_path), monkeypatch, k, limit)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
def test_prefix_of_more_than_16_mi_kmers_vs_reference(tmp_path):
    """The stream writer's block limit at its real size, against the reference binary."""
    uc.case_stream_block_real_limit(tmp_path)

"""GPU parity tests of `meryl union-sum` through the C ABI (mb_union_sum) and the command line: byte-identical with
the golden union dbs written by the unmodified reference binary, with the C oracle on seeded inputs, and -- when
oracle/_ref/meryl travelled to this box -- with the reference binary itself."""
import os
import subprocess

import pytest

import util
import meryl_b200 as mb
from meryl_b200 import synth
import union_cases as uc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", util.union_golden_cases())
def test_union_golden(tmp_path, name):
    uc.case_golden(str(tmp_path), name)


@pytest.mark.parametrize("k", [7, 8, 15, 21, 31, 32, 33, 51, 64])                # (k = 6: tests/test_gpu_zzzz_late.py)
def test_union_k_sweep(tmp_path, k):
    uc.case_k_sweep(str(tmp_path), k, n=60_000)


def test_union_of_unions_empty_single(tmp_path):
    uc.case_union_of_unions(str(tmp_path), n=40_000)


def test_union_low_complexity(tmp_path):
    uc.case_low_complexity(str(tmp_path))


def test_union_groups(tmp_path, monkeypatch):
    uc.case_groups(str(tmp_path), monkeypatch)


def test_union_cli_and_reference(tmp_path):
    """Count two read sets with the GPU count path, union-sum them with the CLI; the result must equal the oracle's
    (and the reference binary's) union of the oracle's counts."""
    k = 21
    fa, fb = tmp_path / "a.fq", tmp_path / "b.fq"
    fa.write_bytes(synth.reads_fastq_bytes(300_000, 30_000, 150, 0.002, 5))
    fb.write_bytes(synth.reads_fastq_bytes(300_000, 30_000, 150, 0.002, 6))
    ga, gb, gu = str(tmp_path / "ga.meryl"), str(tmp_path / "gb.meryl"), str(tmp_path / "gu.meryl")
    mb.count([str(fa)], ga, k, memory_gb=2)
    mb.count([str(fb)], gb, k, memory_gb=2)
    p = subprocess.run([mb.CLI_PATH, "union-sum", ga, gb, "output", gu], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    ou = str(tmp_path / "ou.meryl")
    util.run_oracle(["union-sum", ga, gb, "output", ou])
    util.assert_same_db(ou, gu)
    if util.have_ref():
        ru = str(tmp_path / "ru.meryl")
        util.run_ref(["threads=8", "union-sum", ga, gb, "output", ru])
        util.assert_same_db(ru, gu)


def test_union_properties_large():
    """Size-independent checks at a size the oracle does not reach quickly: union-sum of a db with itself doubles every
    count and keeps the k-mers; totals add up."""
    import tempfile
    k = 21
    with tempfile.TemporaryDirectory() as t:
        fq = os.path.join(t, "r.fq")
        with open(fq, "wb") as fh:
            fh.write(synth.reads_fastq_bytes(5_000_000, 400_000, 150, 0.001, 11))
        a, u = os.path.join(t, "a.meryl"), os.path.join(t, "u.meryl")
        sa = mb.count([fq], a, k, memory_gb=4)
        st = mb.union_sum([a, a], u)
        assert st.nInputKmers == 2 * sa.nDistinct and st.nOutputKmers == sa.nDistinct
        assert st.nTotal == 2 * sa.nTotal and st.nUnique == 0 and st.nDistinct == sa.nDistinct


@pytest.mark.parametrize("op", mb.SET_OPERATIONS)
def test_set_operations(tmp_path, op):
    uc.case_set_operations(str(tmp_path), op, n=40_000)


def test_value_operations(tmp_path):
    uc.case_value_operations(str(tmp_path), n=40_000)


def test_union_small_ring(tmp_path, monkeypatch):
    uc.case_small_ring(str(tmp_path), monkeypatch)


@pytest.mark.parametrize("k,decoder", [(21, "warp"), (21, "thread"), (51, "warp"), (64, "thread"), (32, "warp"), (31, "auto")])
def test_union_long_blocks(tmp_path, monkeypatch, k, decoder):
    uc.case_long_blocks(str(tmp_path), monkeypatch, k, decoder)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
def test_threshold_words_vs_reference(tmp_path):
    uc.case_threshold_words(str(tmp_path), n=60_000)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
def test_cli_operation_tree_vs_reference(tmp_path):
    """`meryl-b200 union-sum [ count a ] [ greater-than 1 [ count b ] ] output x`: sub-operations run first into temporary
    databases; the directory must be the one the reference writes for the same command line."""
    a, b = tmp_path / "a.fa", tmp_path / "b.fq"
    a.write_bytes(synth.fasta_bytes(3, 40_000, 5))
    b.write_bytes(synth.reads_fastq_bytes(40_000, 3_000, 150, 0.002, 6))
    def words(tag):                                            # (the reference wants an output for every count, in a tree too)
        return ["k=21", "memory=2", "union-sum", "[", "count", str(a), "output", str(tmp_path / (tag + "_a.meryl")), "]",
                "[", "greater-than", "1", "[", "count", str(b), "output", str(tmp_path / (tag + "_b.meryl")), "]", "]"]
    ref, out = str(tmp_path / "ref.meryl"), str(tmp_path / "cli.meryl")
    p = subprocess.run([util.REF_BIN] + words("ref") + ["output", ref], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    p = subprocess.run([mb.CLI_PATH] + words("cli") + ["output", out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    util.assert_same_db(ref, out)
    util.assert_same_db(str(tmp_path / "ref_a.meryl"), str(tmp_path / "cli_a.meryl"))
    assert not [n for n in os.listdir(str(tmp_path)) if ".tmp-" in n]          # the temporary databases are gone


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
def test_count_suffix_vs_reference(tmp_path):
    uc.case_count_suffix(str(tmp_path), n=30_000, full=True)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
def test_cli_count_suffix(tmp_path):
    fq = tmp_path / "r.fq"
    fq.write_bytes(synth.reads_fastq_bytes(30_000, 2_000, 150, 0.003, 8))
    ref, out = str(tmp_path / "ref.meryl"), str(tmp_path / "cli.meryl")
    words = ["k=14", "memory=2", "count-suffix=AC", "count", str(fq), "output"]
    assert subprocess.run([util.REF_BIN] + words + [ref], capture_output=True).returncode == 0
    p = subprocess.run([mb.CLI_PATH] + words + [out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    util.assert_same_db(ref, out)
    assert not os.path.exists(out + ".tmp-count")


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
@pytest.mark.parametrize("k", [21, 40, 9])
def test_lookup_vs_reference_print(tmp_path, monkeypatch, k):
    uc.case_lookup(str(tmp_path), k, n=20_000, monkeypatch=monkeypatch)

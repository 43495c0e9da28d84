"""`meryl union-sum` through the product's CUDA sources run in the cusim host emulator (tests/cusim): block decode,
merge-path union with summed counts, stream-mode encoder -- byte-compared with the reference's golden union dbs and
with the oracle.  (The GPU versions of the same cases: tests/test_gpu_union.py.)"""
import os
import subprocess

import pytest

import util
import meryl_b200 as mb
import union_cases as uc
from meryl_b200 import synth

SIM_DIR = os.path.join(util.ROOT, "tests", "cusim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libmeryl_b200_sim.so")


@pytest.fixture(scope="module", autouse=True)
def sim_library():
    subprocess.run(["make", "-s", "-j8", "-C", SIM_DIR], check=True)
    saved = (mb.LIB_PATH, mb._lib)
    mb.LIB_PATH, mb._lib = SIM_LIB, None
    yield
    mb.LIB_PATH, mb._lib = saved


@pytest.mark.parametrize("name", util.union_golden_cases())
def test_sim_union_golden(tmp_path, name):
    uc.case_golden(str(tmp_path), name)


@pytest.mark.parametrize("k", [6, 7, 8, 21, 31, 32, 33, 51, 64])       # (k = 6: the stream writer's 12-bit prefix is the whole k-mer)
def test_sim_union_k_sweep(tmp_path, k):
    uc.case_k_sweep(str(tmp_path), k, n=8_000)


def test_sim_union_of_unions_empty_single(tmp_path):
    uc.case_union_of_unions(str(tmp_path), n=6_000)


def test_sim_union_low_complexity(tmp_path):
    uc.case_low_complexity(str(tmp_path))


def test_sim_union_groups(tmp_path, monkeypatch):
    uc.case_groups(str(tmp_path), monkeypatch)


@pytest.mark.parametrize("op", mb.SET_OPERATIONS)
def test_sim_set_operations(tmp_path, op):
    uc.case_set_operations(str(tmp_path), op, n=6_000)


def test_sim_value_operations(tmp_path):
    uc.case_value_operations(str(tmp_path), n=5_000)


def test_sim_union_small_ring(tmp_path, monkeypatch):
    uc.case_small_ring(str(tmp_path), monkeypatch)


@pytest.mark.parametrize("k,decoder", [(21, "warp"), (21, "thread"), (51, "warp"), (64, "thread"), (32, "warp"), (31, "auto")])
def test_sim_union_long_blocks(tmp_path, monkeypatch, k, decoder):
    uc.case_long_blocks(str(tmp_path), monkeypatch, k, decoder)


@pytest.mark.parametrize("k,limit", [(21, 30), (40, 33), (32, 50)])
def test_sim_stream_block_split(tmp_path, monkeypatch, k, limit):
    uc.case_stream_block_split(str(tmp_path), monkeypatch, k, limit)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
def test_sim_stream_block_real_limit_vs_reference(tmp_path):
    uc.case_stream_block_real_limit(tmp_path)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
def test_sim_threshold_words_vs_reference(tmp_path):
    uc.case_threshold_words(str(tmp_path))


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
def test_sim_count_suffix_vs_reference(tmp_path):
    uc.case_count_suffix(str(tmp_path))


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
@pytest.mark.parametrize("k", [21, 40, 9])
def test_sim_lookup_vs_reference_print(tmp_path, monkeypatch, k):
    uc.case_lookup(str(tmp_path), k, monkeypatch=monkeypatch)


def test_sim_new_entry_points_refuse_bad_arguments(tmp_path):
    a = uc.oracle_db(str(tmp_path), "a", synth.fasta_bytes(1, 3_000, 4), 21)
    out = str(tmp_path / "o.meryl")
    for suffix, msg in (("ACGN", "not made of ACGT"), ("", "1..32 bases"), ("A" * 21, "shorter than k"), ("A" * 33, "1..32 bases")):
        with pytest.raises(mb.MerylB200Error) as e:
            mb.keep_suffix(a, suffix, out)
        assert msg in str(e.value)
    b = uc.oracle_db(str(tmp_path), "b", synth.fasta_bytes(1, 3_000, 4), 40)
    with pytest.raises(mb.MerylB200Error) as e:
        mb.keep_suffix(b, "ACG", out)                           # 2k - 2s = 74 > 42: the reference cannot do it either
    assert "not possible" in str(e.value)
    import numpy as np
    with pytest.raises(mb.MerylB200Error) as e:
        mb.lookup(b, np.zeros(3, dtype=np.uint64))                # k = 40 needs the high words
    assert "high words" in str(e.value)
    assert mb.lookup(a, np.zeros(0, dtype=np.uint64)).size == 0
    e_db = uc.oracle_db(str(tmp_path), "e", b">x\nACGT\n", 21)     # a database without k-mers
    assert mb.db_threshold(e_db, distinct=0.5) in (0, 2 ** 64 - 1)
    assert list(mb.lookup(e_db, np.array([5, 6], dtype=np.uint64))) == [0, 0]


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
@pytest.mark.parametrize("k", [21, 40])
def test_sim_cli_print_vs_reference(tmp_path, k):
    """The command line tool with the emulated library in front of the CUDA one (LD_PRELOAD: same C ABI)."""
    uc.case_cli_print(str(tmp_path), k, env=dict(os.environ, LD_PRELOAD=SIM_LIB))


@pytest.mark.skipif(not util.have_ref_lookup(), reason="reference meryl-lookup not built")
@pytest.mark.parametrize("k", [21, 40, 9])
def test_sim_existence_vs_reference_lookup(tmp_path, monkeypatch, k):
    uc.case_existence(str(tmp_path), k, monkeypatch=monkeypatch)


@pytest.mark.skipif(not util.have_ref_lookup(), reason="reference meryl-lookup not built")
@pytest.mark.parametrize("k", [21])
def test_sim_cli_lookup_vs_reference(tmp_path, k):
    uc.case_cli_lookup(str(tmp_path), k, env=dict(os.environ, LD_PRELOAD=SIM_LIB))


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/meryl not built")
def test_sim_cli_print_of_an_operation_tree(tmp_path):
    """`meryl print [ union-sum a b ]` (merylCommandBuilder.C:441-563): the sub-operation runs into a temporary database and is
    printed from that -- the reference binary's lines (its order depends on its threads).  `histogram` / `statistics` of a
    tree: the reference refuses ("can only accept databases"), and so does the tool."""
    import subprocess
    from meryl_b200 import synth
    env = dict(os.environ, LD_PRELOAD=SIM_LIB, TMPDIR=str(tmp_path))
    dbs = []
    for i in range(2):
        fq = tmp_path / ("r%d.fq" % i)
        fq.write_bytes(synth.reads_fastq_bytes(3_000, 400, 120, 0.01, 40 + i))
        db = str(tmp_path / ("d%d.meryl" % i))
        util.run_ref(["k=17", "memory=2", "threads=4", "count", str(fq), "output", db])
        dbs.append(db)
    cli = os.path.join(util.ROOT, "meryl_b200", "meryl-b200")
    tree = ["print", "[", "union-sum", dbs[0], dbs[1], "]"]
    ref = subprocess.run([util.REF_BIN, "threads=4"] + tree, capture_output=True, text=True)
    got = subprocess.run([cli] + tree, capture_output=True, text=True, env=env)
    assert ref.returncode == 0 and got.returncode == 0, (ref.stderr[-500:], got.stderr[-500:])
    assert sorted(ref.stdout.splitlines()) == sorted(got.stdout.splitlines()) and got.stdout
    assert not [f for f in os.listdir(tmp_path) if f.startswith("meryl-b200-")], "the temporary database is removed"
    for word in ("histogram", "statistics"):
        tree = [word, "[", "union-sum", dbs[0], dbs[1], "]"]
        assert subprocess.run([util.REF_BIN, "threads=4"] + tree, capture_output=True).returncode != 0
        assert subprocess.run([cli] + tree, capture_output=True, env=env).returncode != 0

"""`meryl union-sum`: the C oracle's restatement of the reader / merge / stream writer
(oracle/meryl_oracle.c, sections R1-R3, M1, S1) pinned to the UNMODIFIED reference binary -- against the
committed golden dbs it wrote (tests/golden/make_golden_union.py), and, where oracle/_ref/meryl is present,
against the binary itself on more shapes."""
import os

import pytest

import util
from meryl_b200 import synth


@pytest.mark.parametrize("name", util.union_golden_cases())
def test_oracle_union_sum_matches_golden(tmp_path, name):
    dbs, gold = util.load_union_golden(name, str(tmp_path))
    out = str(tmp_path / "ora.meryl")
    util.run_oracle(["union-sum", *dbs, "output", out])
    util.assert_same_db(gold, out)


def _count_ref(tmp_path, name, data, k, words=("memory=2",), op="count"):
    p = tmp_path / (name + ".fa")
    p.write_bytes(data)
    db = str(tmp_path / (name + ".meryl"))
    util.run_ref(["k=%d" % k, *words, "threads=2", op, str(p), "output", db])
    return db


needs_ref = pytest.mark.skipif(not util.have_ref(), reason="reference binary oracle/_ref/meryl not built")


@needs_ref
@pytest.mark.parametrize("k", [6, 8, 21, 33, 64])        # (k = 6: no suffix bits under the 12-bit prefix; the reference asserts for k = 5)
def test_oracle_union_sum_vs_reference(tmp_path, k):
    a = _count_ref(tmp_path, "a", synth.fasta_bytes(2, 40_000, 10 + k), k)
    b = _count_ref(tmp_path, "b", synth.fasta_bytes(1, 40_000, 10 + k) + synth.fasta_bytes(2, 30_000, 99 + k), k)
    ref, ora = str(tmp_path / "ref.meryl"), str(tmp_path / "ora.meryl")
    util.run_ref(["threads=3", "union-sum", a, b, "output", ref])
    util.run_oracle(["union-sum", a, b, "output", ora])
    util.assert_same_db(ref, ora)


@needs_ref
def test_oracle_union_of_unions_and_empty(tmp_path):
    """Stream-written inputs (prefix 12, absent blocks), an input without k-mers, a single input."""
    k = 21
    a = _count_ref(tmp_path, "a", synth.reads_fastq_bytes(30_000, 2_000, 150, 0.01, 1), k)
    b = _count_ref(tmp_path, "b", synth.reads_fastq_bytes(30_000, 2_000, 150, 0.01, 2), k, op="count-forward")
    e = _count_ref(tmp_path, "e", b">x\nACGT\n", k)
    u1, u2 = str(tmp_path / "u1.meryl"), str(tmp_path / "u2.meryl")
    util.run_ref(["union-sum", a, b, "output", u1])
    util.run_ref(["union-sum", e, b, "output", u2])
    for i, ins in enumerate([[u1, u2], [u1, a, e], [e, e]]):
        ref, ora = str(tmp_path / ("ref%d.meryl" % i)), str(tmp_path / ("ora%d.meryl" % i))
        util.run_ref(["union-sum", *ins, "output", ref])
        util.run_oracle(["union-sum", *ins, "output", ora])
        util.assert_same_db(ref, ora)


@needs_ref
def test_oracle_union_sum_low_complexity_big_blocks(tmp_path):
    """Few distinct k-mers with large counts; sums of counts; k-mers sharing long prefixes (small unary deltas)."""
    k = 21
    rep = b">r\n" + b"ACGTTGCA" * 5_000 + b"\n>p\n" + b"A" * 60_000 + b"\n"
    a = _count_ref(tmp_path, "a", rep + synth.fasta_bytes(1, 50_000, 3), k)
    b = _count_ref(tmp_path, "b", rep + rep, k)
    ref, ora = str(tmp_path / "ref.meryl"), str(tmp_path / "ora.meryl")
    util.run_ref(["union-sum", a, b, "output", ref])
    util.run_oracle(["union-sum", a, b, "output", ora])
    util.assert_same_db(ref, ora)


SET_OPS = ["union", "union-min", "union-max", "union-sum", "intersect", "intersect-min", "intersect-max", "intersect-sum",
           "subtract", "difference", "symmetric-difference"]


@pytest.fixture(scope="module")
def three_dbs(tmp_path_factory):
    """Three count dbs (written by the reference binary) that share some k-mers with different counts."""
    t = tmp_path_factory.mktemp("setops")
    k = 21
    common = synth.reads_fastq_bytes(20_000, 1_500, 150, 0.0, 1)
    a = _count_ref(t, "a", common + synth.reads_fastq_bytes(20_000, 700, 150, 0.0, 2), k)
    b = _count_ref(t, "b", common[: len(common) // 2] + synth.reads_fastq_bytes(20_000, 900, 150, 0.0, 3), k)
    c = _count_ref(t, "c", common + common + synth.reads_fastq_bytes(20_000, 300, 150, 0.0, 2), k)
    return a, b, c


@needs_ref
@pytest.mark.parametrize("op", SET_OPS)
def test_oracle_set_operations_vs_reference(tmp_path, three_dbs, op):
    """Every operator of the merge family (merylOp-nextMer.C:559-612): one, two and three inputs, counts that differ."""
    a, b, c = three_dbs
    cases = [("3", [a, b, c])]                             # (every reference run costs ~2.4 s of start-up)
    if op in ("subtract", "difference", "intersect", "symmetric-difference"):
        cases = [("rev", [c, a])] + (cases if op != "symmetric-difference" else [])
    if op in ("union", "subtract"):
        cases.append(("1", [c]))
    for tag, ins in cases:
        ref, ora = str(tmp_path / ("ref" + tag)), str(tmp_path / ("ora" + tag))
        util.run_ref(["threads=3", op, *ins, "output", ref])
        util.run_oracle([op, *ins, "output", ora])
        util.assert_same_db(ref, ora)


@needs_ref
def test_oracle_value_operations_vs_reference(tmp_path, three_dbs):
    """The one-input filters / arithmetic with their overflow and zero corners (merylOp-nextMer.C:497-557)."""
    import union_cases as uc
    a, b, c = three_dbs
    u = str(tmp_path / "u.meryl")
    util.run_ref(["union-sum", a, b, c, "output", u])          # values 1..4
    corners = [("less-than", 3), ("increase", 4294967295), ("decrease", 1), ("multiply", 4294967296), ("divide-round", 3), ("modulo", 2)]
    # (the other operators and corners were checked against the reference binary once, by hand: all sixteen VALUE_CASES agree)
    for i, (op, param) in enumerate(corners):
        ref, ora = str(tmp_path / ("ref%d" % i)), str(tmp_path / ("ora%d" % i))
        util.run_ref(["threads=2", op, str(param), u, "output", ref])
        util.run_oracle([op, str(param), u, "output", ora])
        util.assert_same_db(ref, ora)

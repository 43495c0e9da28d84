"""Pins the C oracle (oracle/meryl_oracle.c) to the UNMODIFIED reference binary
(oracle/_ref/meryl): same seeded input -> the 129 db files must be byte-identical.

Edge cases follow SURVEY.md Appendix A.5.  These tests need oracle/_ref/meryl
(built from /root/reference by oracle/build_ref.sh, or shipped prebuilt to the
GPU box); they are skipped when it is absent -- tests/test_golden.py then still
pins the oracle against committed reference outputs."""
import os

import numpy as np
import pytest

import util
from meryl_b200 import synth

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="reference binary oracle/_ref/meryl not built")


def _both(tmp_path, files: dict[str, bytes], k: int, op: str = "count", memory: str = "2", extra=()):
    names = []
    for name, data in files.items():
        p = tmp_path / name
        p.write_bytes(data)
        names.append(str(p))
    ref, ora = str(tmp_path / "ref.meryl"), str(tmp_path / "ora.meryl")
    args = ["k=%d" % k, "memory=%s" % memory, *extra, op, *names]
    util.run_ref(args + ["threads=1", "output", ref])
    util.run_oracle(args + ["output", ora])
    util.assert_same_db(ref, ora)
    return ref


def test_c1_shape_small(tmp_path):
    _both(tmp_path, {"a.fa": synth.fasta_bytes(4, 100_000, 42)}, 21)


@pytest.mark.parametrize("k", [3, 4, 6, 15, 16, 21, 22, 31, 32, 33, 51, 64])       # (k = 3: simple mode with prefix 6 and no suffix bits)
def test_k_sweep(tmp_path, k):
    _both(tmp_path, {"a.fa": synth.fasta_bytes(2, 30_000, 100 + k)}, k)


@pytest.mark.parametrize("op", ["count-forward", "count-reverse"])
def test_forward_reverse(tmp_path, op):
    _both(tmp_path, {"a.fa": synth.fasta_bytes(2, 40_000, 5)}, 21, op=op)


def test_fastq_reads_with_errors(tmp_path):
    _both(tmp_path, {"r.fq": synth.reads_fastq_bytes(50_000, 8_000, 150, 0.002, 9)}, 21)


def test_fastq_k31_high_coverage(tmp_path):
    _both(tmp_path, {"r.fq": synth.reads_fastq_bytes(5_000, 6_000, 150, 0.0, 3)}, 31)


def test_multiple_inputs_and_formats(tmp_path):
    _both(tmp_path, {"a.fa": synth.fasta_bytes(3, 20_000, 1),
                     "b.fq": synth.reads_fastq_bytes(20_000, 1_000, 100, 0.01, 2),
                     "c.fa": synth.fasta_bytes(1, 50_000, 3, width=60, eol=b"\r\n")}, 21)


def test_parser_edge_cases(tmp_path):
    fa = (b"\n\n>rec1 lower and N\nacgtacgtNNNNacgtacgtacgtacgtacgtacgtac\nACGTRYKMacgtacgtacgtacgtacgtacgtacgtacgt\n"
          b">short\nACGT\n"
          b">empty\n"
          b">after_empty\nACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n"
          b">spaces\nACGT ACGT\tACGTACGT ACGTACGTACGTACGTACGTACGTAC\n"
          b">crlf\r\nACGTTGCAACGTTGCAACGTTGCAACGTTGCA\r\nACGTTGCAACGTTGCAACGTTGCAACGTTGCA\r\n"
          b">palindrome\nACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n"
          b">noeol\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTGGGGGGGGGGGGGGGGGGGGGGGGG")
    fq = (b"@q1\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\n@@@@IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
          b"@q2\nGGGGGGGGGGGGGGGGGGGGGGCCCCCCCCCCCCCCCCCCCCCC\n+q2\n>>>>IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
          b"@q3\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTAAAA\n+\n++++IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n")
    _both(tmp_path, {"e.fa": fa, "e.fq": fq}, 8)
    _both(tmp_path, {"e2.fa": fa, "e2.fq": fq}, 21)


def test_record_longer_than_work_buffer(tmp_path):
    # one 5 Mbp record (80-col) crosses the reference's 2 MB work buffers several times
    _both(tmp_path, {"long.fa": synth.fasta_bytes(1, 5_000_000, 77)}, 21, memory="4")


def test_single_line_long_record(tmp_path):
    s = synth.random_bases(3_000_000, np.random.default_rng(11)).tobytes()
    _both(tmp_path, {"oneline.fa": b">x\n" + s + b"\n"}, 25, memory="4")


def test_low_complexity_big_counts_and_many_subblocks(tmp_path):
    # few distinct k-mers with large counts; one giant block -> > 32 sub-blocks? and histogram with many values
    rng = np.random.default_rng(5)
    unit = synth.random_bases(3000, rng).tobytes()
    recs = []
    for i in range(400):
        n = 1 + (i * 7) % 300
        recs.append(b">r%d\n" % i + unit[: 200 + n * 9] + b"\n")
    recs.append(b">polyA\n" + b"A" * 100_000 + b"\n")
    _both(tmp_path, {"rep.fa": b"".join(recs)}, 21)


def test_empty_and_no_kmers(tmp_path):
    _both(tmp_path, {"e.fa": b">only\nACGT\n"}, 21)


@pytest.mark.parametrize("k,n,mem", [(21, 10_125_060, 8), (21, 626_888_890, 32), (21, 210_000_000_000, 256),
                                     (21, 210_000_000_000, 1024), (31, 210_000_000_000, 512),
                                     (51, 100_000_000_000, 256), (51, 100_000_000_000, 1024),
                                     (16, 5_000_000, 1), (64, 3_000_000_000, 64), (10, 1000, 1),
                                     (21, 86_670_000_000, 180)])
def test_prefix_width_planner(k, n, mem):
    assert util.oracle_plan(k, n, int(mem * 1024 ** 3)) == util.ref_wprefix(k, n, mem)


@pytest.mark.parametrize("k", [15, 21, 40])
def test_compress_homopolymer_runs(tmp_path, k):
    """`count compress`: runs of one letter collapse (case-insensitive, N runs too) before k-mers are formed."""
    rng = np.random.default_rng(k)
    recs = []
    for i in range(40):
        seq = bytearray()
        for _ in range(400):
            base = b"ACGTacgtN"[int(rng.integers(0, 9))]
            seq += bytes([base]) * int(rng.integers(1, 6))
        recs.append(b">r%d\n" % i + bytes(seq) + b"\n")
    recs.append(b">polyA\n" + b"A" * 5000 + b"C" + b"a" * 70 + b"\n" + b"AAAA\n" + b"aaCC\n")   # runs across line ends
    _both(tmp_path, {"hp.fa": b"".join(recs), "reads.fq": synth.reads_fastq_bytes(20_000, 500, 150, 0.01, 3)}, k, extra=("compress",))

"""CPU-side tests of the product's host code (no GPU needed): the C ABI loads and exports what
include/meryl_b200.h declares, the prefix-width planner agrees with the oracle (which is pinned to
the reference), the FASTA/FASTQ parser emits the oracle's base stream, and device entry points
fail loudly (no CPU fallback) when no GPU exists."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import util
import meryl_b200 as mb
from meryl_b200 import synth

ROOT = util.ROOT


@pytest.fixture(scope="module", autouse=True)
def _build_product():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "meryl_b200", "csrc")], check=True)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "meryl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = ctypes.CDLL(mb.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mb.MerylB200Error) as e:
        mb.Counter(21, 10)
    assert e.value.code == -2
    p = subprocess.run([mb.CLI_PATH, "k=21", "count", os.path.join(ROOT, "README.md"), "output", "/tmp/x.meryl"],
                       capture_output=True, text=True)
    assert p.returncode == 1


@pytest.mark.parametrize("k", [3, 5, 6, 8, 13, 14, 15, 16, 17, 20, 21, 22, 31, 32, 33, 51, 64])
def test_planner_matches_oracle(k):
    rng = np.random.default_rng(k)
    ns = [1, 1000, 10_125_060, 626_888_890, 86_670_000_000, 210_000_000_000] + [int(x) for x in 10 ** rng.uniform(3, 12, 12)]
    mems = [1, 2, 8, 32, 180, 256, 1024] + [float(x) for x in rng.uniform(0.5, 2048, 6)]
    L = util.oracle_lib()
    for n in ns:
        for m in mems:
            mbytes = int(m * 1024 ** 3)
            wp = ctypes.c_uint32(0)
            used = ctypes.c_uint64(0)
            nb = ctypes.c_uint32(0)
            rc = L.ora_plan(k, n, mbytes, ctypes.byref(wp), ctypes.byref(used), ctypes.byref(nb))
            if rc != 0:
                with pytest.raises(mb.MerylB200Error):
                    mb.plan(k, n, mbytes)
                continue
            p = mb.plan(k, n, mbytes)
            assert (p.w_prefix, p.mem_used, p.n_batches) == (wp.value, used.value, nb.value), (k, n, m)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
@pytest.mark.parametrize("k,n,mem", [(21, 10_125_060, 8), (21, 210_000_000_000, 256), (31, 210_000_000_000, 512),
                                     (51, 100_000_000_000, 1024), (16, 5_000_000, 1), (14, 100_000_000, 4)])
def test_planner_matches_reference(k, n, mem):
    assert mb.plan(k, n, int(mem * 1024 ** 3)).w_prefix == util.ref_wprefix(k, n, mem)


def _collapse(s: bytes) -> bytes:
    return re.sub(rb"\.+", b".", s)


def _product_parse(data: bytes) -> bytes:
    L = mb.lib()
    n = ctypes.c_uint64(0)
    buf = ctypes.create_string_buffer(len(data) + 64)
    rc = L.mb_dbg_parse(data, len(data), buf, len(data) + 64, ctypes.byref(n))
    if rc != 0:
        raise mb.MerylB200Error(rc, L.mb_last_error().decode())
    return buf.raw[: n.value]


EDGE = [
    b"",
    b"\n\n",
    b">only header",
    b">h\nACGT",
    b">h\nACGT\n",
    b">h\r\nACGT\r\nTTTT\r\n>i\r\nGG\r\n",
    b">h\rACGT\rTT\r>i\rGG",
    b">a\n>b\nACGT\n",
    b">a\n\n\n>b\n\nAC\n\nGT\n\n",
    b">a\nAC GT\tTT\n",
    b"@r1\nACGT\n+\nIIII\n@r2\nTTTT\n+r2\n@@@@\n",
    b"@r1\nACGT\n+\n>>>>\n@r2\nTTTT\n+\n++++\n",
    b"@r1\nACGT\n+\nIIII",
    b"@r1\nACGT\n+\n",
    b"@r1\nACGT\n+",
    b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGGGG\r\n+\r\nIIII\r\n",
    b">x\nACGT+TT\nAA\n+AC\nGGGG\n>y\nCC\n",
    b">x\nACGT\n+\nAAAA\nCCCC\n",
    b">x\nAC>GT\n@T\n",
    b"  \n>x desc\nNNNNACGTNNNN\n",
]


@pytest.mark.parametrize("i", range(len(EDGE)))
def test_parser_edge_cases_match_oracle(i):
    data = EDGE[i]
    assert _collapse(_product_parse(data)) == _collapse(util.oracle_parse(data))


def test_parser_fuzz_matches_oracle():
    rng = np.random.default_rng(123)
    alphabet = np.frombuffer(b"ACGTacgtN>@+\n\n\r \tIx", dtype=np.uint8)
    for it in range(400):
        n = int(rng.integers(1, 200))
        body = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
        data = (b">" if it % 2 else b">s\n") + body
        assert _collapse(_product_parse(data)) == _collapse(util.oracle_parse(data)), data


def test_parser_rejects_non_sequence_and_sam():
    for data in [b"hello world\n", b"@HD\tVN:1.6\n@SQ\tSN:x\n", b"@r1\tfoo\nACGT\t0\n"]:
        with pytest.raises(mb.MerylB200Error):
            _product_parse(data)
        with pytest.raises(ValueError):
            util.oracle_parse(data)


def test_parser_large_fastq_and_fasta():
    for data in [synth.reads_fastq_bytes(10_000, 3_000, 150, 0.01, 5), synth.fasta_bytes(3, 100_000, 6),
                 synth.fasta_bytes(2, 50_000, 7, width=70, eol=b"\r\n")]:
        assert _collapse(_product_parse(data)) == _collapse(util.oracle_parse(data))


def _product_parse_parallel(data: bytes, block: int) -> bytes:
    L = mb.lib()
    L.mb_dbg_parse_parallel.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint64,
                                        ctypes.POINTER(ctypes.c_uint64)]
    n = ctypes.c_uint64(0)
    buf = ctypes.create_string_buffer(len(data) + 64)
    rc = L.mb_dbg_parse_parallel(data, len(data), block, buf, len(data) + 64, ctypes.byref(n))
    if rc != 0:
        raise mb.MerylB200Error(rc, L.mb_last_error().decode())
    return buf.raw[: n.value]


def test_parallel_parser_equals_sequential(monkeypatch):
    """File inputs are parsed in blocks on several threads, cut at guessed record starts; the guess is verified against the
    state machine, so the result must be byte-identical to one sequential pass -- including the traps: quality lines that
    begin with '@', '>' or '+', CRLF, multi-line FASTA, cuts that fall inside a record, blocks that end mid-line."""
    monkeypatch.setenv("MB_PARSE_THREADS", "5")
    monkeypatch.setenv("MB_PARSE_MIN_PIECE", "64")
    rng = np.random.default_rng(9)
    trap = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"ACGTTGCAAC"[: 3 + i % 8] * 3, [b"@", b">", b"+", b"I"][i % 4] * (3 * (3 + i % 8))) for i in range(400))
    cases = [synth.reads_fastq_bytes(10_000, 2_000, 150, 0.01, 5), synth.fasta_bytes(40, 3_000, 6),
             synth.fasta_bytes(5, 20_000, 7, width=70, eol=b"\r\n"), trap, trap.replace(b"\n", b"\r\n"),
             b">x\n" + b"ACGT" * 50_000 + b"\n"]
    for data in cases:
        want = _product_parse(data)
        for block in (len(data) + 1, 4096, 1000 + int(rng.integers(0, 5000))):
            assert _product_parse_parallel(data, block) == want
    alphabet = np.frombuffer(b"ACGTacgtN>@+\n\n\r \tIx", dtype=np.uint8)
    for it in range(200):                                  # fuzz: arbitrary bytes, every cut guess gets verified or redone
        n = int(rng.integers(200, 3000))
        data = (b">" if it % 2 else b">s\n") + alphabet[rng.integers(0, len(alphabet), n)].tobytes()
        assert _collapse(_product_parse_parallel(data, 700)) == _collapse(_product_parse(data)), data


def test_estimate_kmers_in_file(tmp_path):
    p = tmp_path / "a.fa"
    p.write_bytes(b">x\nACGT\n" * 100)
    assert mb.estimate_kmers_in_file(str(p)) == 800
    g = tmp_path / "a.fa.gz"
    g.write_bytes(b"x" * 1000)
    assert mb.estimate_kmers_in_file(str(g)) == 3000
    b = tmp_path / "a.fa.bz2"
    b.write_bytes(b"x" * 1001)
    assert mb.estimate_kmers_in_file(str(b)) == int(1001 * 3.5)


def _cli(args, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    return subprocess.run([mb.CLI_PATH] + args, capture_output=True, text=True, env=e)


def test_cli_errors_like_the_reference(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">x\nACGTACGTACGTACGTACGTACGTACGT\n")
    p = _cli(["count", str(fa), "output", str(tmp_path / "o.meryl")])
    assert p.returncode == 1 and "Kmer size not supplied" in p.stderr          # merylOp-count.C:311-312
    p = _cli(["k=21", "count", str(fa)])
    assert p.returncode == 1 and "No output specified" in p.stderr             # merylOp-count.C:314-315
    p = _cli(["k=21", "print", "x.meryl"])
    assert p.returncode == 1 and "no meryl database given to print" in p.stderr
    p = _cli(["k=21", "compare", "x.meryl"])
    assert p.returncode == 1 and "not part of `meryl count`" in p.stderr
    p = _cli(["k=21", "intersect", "x.meryl", "y.meryl", "output", "z.meryl"])
    assert p.returncode == 1 and "is not a meryl database" in p.stderr
    p = _cli(["k=21", "union-sum", "x.meryl", "y.meryl", "output", "z.meryl"])
    assert p.returncode == 1 and "is not a meryl database" in p.stderr
    p = _cli(["union-sum", "output", "z.meryl"])
    assert p.returncode == 1 and "needs at least one input database" in p.stderr
    p = _cli(["k=21", "gpus=3", "count", str(fa), "output", str(tmp_path / "o.meryl")])     # (gpus= is ours: one process per GPU)
    assert p.returncode == 1 and "must divide 64" in p.stderr
    p = _cli(["k=21", "gpus=2", "count", str(fa), "output", str(tmp_path / "o.meryl")])     # no GPU here: every rank says so
    assert p.returncode == 1 and "rank" in p.stderr and not (tmp_path / "o.meryl" / "merylIndex").exists()


def test_cli_config_only_prints_the_canu_line(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_bytes(synth.fasta_bytes(1, 10_000, 3))
    p = _cli(["-C", "k=21", "memory=8", "n=626888890", "count", str(fa)])
    assert p.returncode == 0
    assert "Configured complex mode for" in p.stderr and "GB memory per batch, and up to" in p.stderr
    assert "prefix width %d bits" % mb.plan(21, 626_888_890, 8 << 30).w_prefix in p.stderr
    assert not (tmp_path / "o.meryl").exists()


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
def test_cli_hands_other_operations_to_the_reference(tmp_path):
    """With MERYL_REFERENCE set, anything that is not a count is exec'd to the reference unchanged."""
    inp, gold, words, _ = util.load_golden("edge_k21", str(tmp_path))
    p = _cli(["statistics", gold], env={"MERYL_REFERENCE": util.REF_BIN})
    q = subprocess.run([util.REF_BIN, "statistics", gold], capture_output=True, text=True)
    assert p.returncode == q.returncode == 0 and p.stdout == q.stdout and "unique" in p.stdout


def test_cli_operation_trees_become_flat_commands(tmp_path):
    """'[' ... ']' (merylCommandBuilder.C:441-563): a sub-operation runs first, into its own output or a temporary database next
    to the root's output; k= is global, the outermost options are handed down.  MB_CLI_DRY=1 prints the flat command lines."""
    env = dict(os.environ, MB_CLI_DRY="1")
    def dry(words):
        p = subprocess.run([mb.CLI_PATH] + words, capture_output=True, text=True, env=env, cwd=str(tmp_path))
        return p.returncode, p.stdout.splitlines(), p.stderr
    rc, out, _ = dry("k=21 threads=4 union-sum output u.meryl [ count a.fa output a.meryl ] [ count memory=2 b.fa ] c.meryl".split())
    assert rc == 0 and out == ["k=21 threads=4 count a.fa output a.meryl",
                               "k=21 threads=4 count memory=2 b.fa output u.meryl.tmp-0",
                               "k=21 threads=4 union-sum output u.meryl a.meryl u.meryl.tmp-0 c.meryl"]
    rc, out, _ = dry("[ difference [ greater-than 1 [ count k=15 x.fa ] ] y.meryl output d.meryl ]".split())
    assert rc == 0 and out == ["count k=15 x.fa output d.meryl.tmp-0",
                               "k=15 greater-than 1 d.meryl.tmp-0 output d.meryl.tmp-1",
                               "k=15 difference d.meryl.tmp-1 y.meryl output d.meryl"]
    rc, _, err = dry("union [ count k=9 a ]".split())
    assert rc == 1 and "No output specified" in err
    rc, _, err = dry("union [ count k=9 a output z".split())
    assert rc == 1 and "without a ']'" in err


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
def test_cli_histogram_and_statistics_print_what_the_reference_prints(tmp_path):
    """`meryl histogram` / `meryl statistics` (merylOp-histogram.C:20-100) from the master index: no device involved.  On
    every golden count db (k = 8 ... 64: the 'missing' line wraps for k >= 32 like the reference's) and a golden union."""
    dbs = []
    for name in util.golden_cases():
        (tmp_path / name).mkdir()
        dbs.append(util.load_golden(name, str(tmp_path / name))[1])
    (tmp_path / "u").mkdir()
    dbs.append(util.load_union_golden(util.union_golden_cases()[0], str(tmp_path / "u"))[1])
    for db in dbs:
        for word in ("histogram", "statistics"):
            ours = _cli([word, db])
            ref = subprocess.run([util.REF_BIN, word, db], capture_output=True, text=True)
            assert ours.returncode == 0 and ref.returncode == 0, ours.stderr
            assert ours.stdout == ref.stdout and ours.stdout
    p = _cli(["histogram", str(tmp_path)])
    assert p.returncode == 1 and "no meryl database" in p.stderr

"""GPU parity tests of the whole `meryl count` path through the C ABI: FASTA/FASTQ bytes in,
the 129-file db out, compared byte for byte with
  - the committed golden dbs written by the unmodified reference binary (tests/golden/),
  - the C oracle on seeded synthetic inputs (SURVEY.md 8(d) shapes, Appendix A.5 edge cases),
  - the reference binary itself when oracle/_ref/meryl travelled to this box,
and, at sizes the oracle cannot reach quickly, through size-independent properties."""
import os
import subprocess

import numpy as np
import pytest

import util
import meryl_b200 as mb
from meryl_b200 import synth

pytestmark = pytest.mark.gpu

_OPS = {"count": "count", "count-forward": "count-forward", "count-reverse": "count-reverse"}


def _words(words):
    k = mem = None
    op = "count"
    for w in words:
        if w.startswith("k="):
            k = int(w[2:])
        elif w.startswith("memory="):
            mem = float(w[7:])
        elif w in _OPS:
            op = w
    return k, mem, op


def _gpu_db(tmp, paths, k, mem, op, n=None, name="gpu.meryl"):
    out = os.path.join(tmp, name)
    mb.count(paths, out, k, op, memory_gb=mem, n=n)
    return out


@pytest.mark.parametrize("name", util.golden_cases())
def test_golden_db_bit_exact(tmp_path, name):
    inp, gold, words, _ = util.load_golden(name, str(tmp_path))
    k, mem, op = _words(words)
    out = _gpu_db(str(tmp_path), [inp], k, mem, op)
    util.assert_same_db(gold, out)


def test_cli_golden(tmp_path):
    inp, gold, words, _ = util.load_golden("fastq_k31", str(tmp_path))
    out = str(tmp_path / "cli.meryl")
    p = subprocess.run([mb.CLI_PATH] + words + ["threads=4", inp, "output", out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Configured complex mode" in p.stderr
    util.assert_same_db(gold, out)


def _vs_oracle(tmp_path, files, k, op="count", memory="2"):
    paths = []
    for name, data in files.items():
        p = tmp_path / name
        p.write_bytes(data)
        paths.append(str(p))
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=%d" % k, "memory=%s" % memory, op, *paths, "output", ora])
    out = _gpu_db(str(tmp_path), paths, k, float(memory), op)
    util.assert_same_db(ora, out)
    return out


def test_c1_config_10mb_fasta_k21(tmp_path):
    """BASELINE.json configs[0]: k=21 count on 10 x 1 Mbp FASTA (seed 42), bit-exact db."""
    out = _vs_oracle(tmp_path, {"c1.fa": synth.fasta_bytes(10, 1_000_000, 42)}, 21, memory="8")
    if util.have_ref():
        ref = str(tmp_path / "ref.meryl")
        util.run_ref(["k=21", "threads=4", "memory=8", "count", str(tmp_path / "c1.fa"), "output", ref])
        util.assert_same_db(ref, out)


@pytest.mark.parametrize("k", [4, 6, 13, 14, 15, 16, 22, 31, 32, 33, 51, 64])       # (k = 3: tests/test_gpu_zzzz_late.py)
def test_k_sweep(tmp_path, k):
    _vs_oracle(tmp_path, {"a.fa": synth.fasta_bytes(2, 30_000, 100 + k)}, k)


@pytest.mark.parametrize("op", ["count-forward", "count-reverse"])
def test_forward_reverse(tmp_path, op):
    _vs_oracle(tmp_path, {"a.fq": synth.reads_fastq_bytes(20_000, 2_000, 120, 0.01, 5)}, 21, op=op)


def test_reads_with_errors_k21_and_k31(tmp_path):
    data = synth.reads_fastq_bytes(100_000, 30_000, 150, 0.002, 9)
    _vs_oracle(tmp_path, {"r.fq": data}, 21)
    (tmp_path / "b").mkdir()
    _vs_oracle(tmp_path / "b", {"r.fq": data}, 31)


def test_multiple_inputs_and_formats(tmp_path):
    _vs_oracle(tmp_path, {"a.fa": synth.fasta_bytes(3, 20_000, 1),
                          "b.fq": synth.reads_fastq_bytes(20_000, 1_000, 100, 0.01, 2),
                          "c.fa": synth.fasta_bytes(1, 50_000, 3, width=60, eol=b"\r\n")}, 21)


def test_low_complexity_big_counts(tmp_path):
    rng = np.random.default_rng(5)
    unit = synth.random_bases(3000, rng).tobytes()
    recs = [b">r%d\n" % i + unit[: 200 + (1 + (i * 7) % 300) * 9] + b"\n" for i in range(400)]
    recs.append(b">polyA\n" + b"A" * 100_000 + b"\n")
    recs.append(b">polyAC\n" + b"AC" * 80_000 + b"\n")
    _vs_oracle(tmp_path, {"rep.fa": b"".join(recs)}, 21)


def test_no_kmers_at_all(tmp_path):
    _vs_oracle(tmp_path, {"e.fa": b">only\nACGT\n"}, 21)


def test_gz_input(tmp_path):
    import gzip
    data = synth.fasta_bytes(2, 20_000, 8)
    plain = tmp_path / "a.fa"
    plain.write_bytes(data)
    gz = tmp_path / "b.fa.gz"
    gz.write_bytes(gzip.compress(data))
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=21", "wprefix=10", "count", str(plain), "output", ora])
    out = str(tmp_path / "gpu.meryl")
    with mb.Counter(21, 10) as c:
        c.add_file(str(gz))
        c.count()
        c.write_db(out)
    util.assert_same_db(ora, out)


def test_add_bases_seam_equals_file_path(tmp_path):
    """merylInput::loadBases seam: pushing base chunks == parsing the file."""
    data = synth.reads_fastq_bytes(30_000, 4_000, 150, 0.01, 21)
    stream = util.oracle_parse(data)
    a, b = str(tmp_path / "a.meryl"), str(tmp_path / "b.meryl")
    with mb.Counter(21, 10) as c:
        c.add_file_bytes(data)
        c.count()
        c.write_db(a)
    with mb.Counter(21, 10) as c:
        for rec in stream.split(b"."):
            if rec:
                h = len(rec) // 2
                c.add_bases(rec[:h], end_of_sequence=False)     # a record arrives in pieces
                c.add_bases(rec[h:], end_of_sequence=True)
        c.count()
        c.write_db(b)
    util.assert_same_db(a, b)


def test_counted_arrays_and_histogram_properties():
    """Size-independent checks on a larger input: sum of counts == instances, strictly ascending
    distinct k-mers, histogram totals consistent with the master index statistics."""
    nreads, readlen, k = 400_000, 150, 21
    data = synth.reads_fastq_bytes(2_000_000, nreads, readlen, 0.001, 31)
    with mb.Counter(k, 12, flags=mb.FLAG_KEEP_COUNTED) as c:
        c.add_file_bytes(data)
        c.count()
        st = c.stats()
        lo, hi, cnt = c.counted()
        v, o = c.histogram()
    assert st.nKmers == nreads * (readlen - k + 1)
    assert int(cnt.sum(dtype=np.uint64)) == st.nKmers == st.nTotal
    assert len(lo) == st.nDistinct and np.all(lo[1:] > lo[:-1])
    assert int(o.sum()) == st.nDistinct and int((v * o).sum()) == st.nTotal
    assert int(o[v == 1].sum()) == st.nUnique == int((cnt == 1).sum())
    # canonical: no k-mer is larger than its reverse complement
    sample = lo[:: max(1, len(lo) // 2000)]
    for x in sample[:200]:
        x = int(x)
        rc = 0
        for i in range(k):
            rc |= (((x >> (2 * i)) & 3) ^ 2) << (2 * (k - 1 - i))
        assert x <= rc


def test_multi_pass_equals_single_pass(tmp_path):
    """A small workspace forces several prefix-range passes; the db must not change."""
    data = synth.reads_fastq_bytes(300_000, 400_000, 150, 0.002, 77)
    a, b = str(tmp_path / "a.meryl"), str(tmp_path / "b.meryl")
    with mb.Counter(21, 12) as c:
        c.add_file_bytes(data)
        c.count()
        one = c.stats().nPasses
        c.write_db(a)
    with mb.Counter(21, 12, workspace_bytes=1 << 28) as c:
        c.add_file_bytes(data)
        c.count()
        many = c.stats().nPasses
        c.write_db(b)
    assert one == 1 and many > 1, (one, many)
    util.assert_same_db(a, b)


@pytest.mark.parametrize("k,op", [(21, "count"), (21, "count-forward"), (22, "count-reverse"), (31, "count"), (51, "count"), (64, "count")])
@pytest.mark.parametrize("sort_version", ["2", "3"])
def test_multi_pass_file_aligned_prefilter(tmp_path, monkeypatch, k, op, sort_version):
    """Capped passes (MB_MAX_PASS_KEYS) exercise the file-aligned pass plan and the extraction kernel's bit-sliced
    pass-membership filter; run with both sweep kernels (MB_SORT is read when the first sort is planned, so each
    combination runs in its own process)."""
    data = synth.reads_fastq_bytes(200_000, 60_000, 150, 0.003, 100 + k)
    fq = tmp_path / "r.fq"
    fq.write_bytes(data)
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "gpu.meryl")
    util.run_oracle(["k=%d" % k, "memory=2", op, str(fq), "output", ora])
    env = dict(os.environ, MB_MAX_PASS_KEYS="1500000", MB_SORT=sort_version)
    p = subprocess.run([mb.CLI_PATH, "k=%d" % k, "memory=2", op, str(fq), "output", out], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    import re
    assert int(re.search(r"in (\d+) pass", p.stderr).group(1)) >= 3, p.stderr
    util.assert_same_db(ora, out)


@pytest.mark.parametrize("cap", [0, 1_500_000])
def test_streamed_output_dir(tmp_path, monkeypatch, cap):
    """mb_set_output_dir: the data files go from the device to the directory while the count runs (csrc/filesink.cpp, one
    worker and CUDA stream per file); write_db adds the block indexes and merylIndex.  Same 129 files as the oracle."""
    if cap:
        monkeypatch.setenv("MB_MAX_PASS_KEYS", str(cap))
    fq = tmp_path / "r.fq"
    fq.write_bytes(synth.reads_fastq_bytes(300_000, 60_000, 150, 0.004, 31))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "gpu.meryl")
    util.run_oracle(["k=21", "wprefix=14", "count", str(fq), "output", ora])
    with mb.Counter(21, 14) as c:
        c.set_output_dir(out)
        c.add_file(str(fq))
        c.count()
        assert (c.stats().nPasses > 1) == bool(cap)
        with pytest.raises(mb.MerylB200Error):
            c.image(0)
        c.write_db(out)
    util.assert_same_db(ora, out)


def test_compress_cli(tmp_path):
    """`meryl-b200 k=K count compress <files>`: homopolymer-compressed counting, against the oracle (and the reference)."""
    rng = np.random.default_rng(8)
    seq = bytearray()
    for _ in range(60_000):
        seq += bytes([b"ACGTacgtN"[int(rng.integers(0, 9))]]) * int(rng.integers(1, 5))
    fa = tmp_path / "hp.fa"
    fa.write_bytes(b">a\n" + bytes(seq) + b"\n")
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "gpu.meryl")
    util.run_oracle(["k=21", "memory=2", "compress", "count", str(fa), "output", ora])
    p = subprocess.run([mb.CLI_PATH, "k=21", "memory=2", "count", "compress", str(fa), "output", out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    util.assert_same_db(ora, out)
    if util.have_ref():
        ref = str(tmp_path / "ref.meryl")
        util.run_ref(["k=21", "memory=2", "threads=4", "count", "compress", str(fa), "output", ref])
        util.assert_same_db(ref, out)


def test_device_generator_and_device_ingest_match_oracle(tmp_path):
    """bench.py's input path: reads generated on the device (csrc/synth.cu) and ingested from device
    memory must give the same db as the oracle run on the same reads written as a FASTQ file."""
    import torch
    n_reads, L, G, err, seed = 30_000, 150, 500_000, 0.004, 11
    rec0, rec1 = mb.synth_record_bytes(L, 0), mb.synth_record_bytes(L, 1)
    assert (rec0, rec1) == (151, 319)
    d0 = torch.empty(n_reads * rec0, dtype=torch.uint8, device="cuda")
    d1 = torch.empty(n_reads * rec1, dtype=torch.uint8, device="cuda")
    mb.synth_reads_device(d0.data_ptr(), 5, n_reads, L, G, err, seed, 0)
    mb.synth_reads_device(d1.data_ptr(), 5, n_reads, L, G, err, seed, 1)
    torch.cuda.synchronize()
    fq = d1.cpu().numpy().tobytes()
    stream = d0.cpu().numpy().tobytes()
    assert fq.startswith(b"@r000000000005\n") and fq.count(b"\n") == 4 * n_reads
    assert util.oracle_parse(fq).replace(b"..", b".") == stream + b"."[:0] or util.oracle_parse(fq).rstrip(b".") == stream.rstrip(b".")
    p = tmp_path / "s.fq"
    p.write_bytes(fq)
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=21", "wprefix=12", "count", str(p), "output", ora])
    out = str(tmp_path / "gpu.meryl")
    with mb.Counter(21, 12) as c:
        half = (n_reads // 2) * rec0
        c.add_bases_device(d0.data_ptr(), half)                       # two device chunks of whole records
        c.add_bases_device(d0.data_ptr() + half, n_reads * rec0 - half) if half % 16 == 0 else c.add_bases(stream[half:], True)
        c.count()
        c.write_db(out)
    util.assert_same_db(ora, out)


def test_pinned_host_ingest_fast_path(tmp_path):
    """mb_add_bases on a large page-locked buffer takes the in-place DMA path (64 MB pieces that cut
    records anywhere); the db must equal the one from the ordinary staged path."""
    import torch
    n_reads, L = 2_000_000, 150                                       # 302 MB of base stream: > 4 staging buffers
    rec0 = mb.synth_record_bytes(L, 0)
    d0 = torch.empty(n_reads * rec0, dtype=torch.uint8, device="cuda")
    mb.synth_reads_device(d0.data_ptr(), 0, n_reads, L, 3_000_000, 0.001, 4, 0)
    host = torch.empty(n_reads * rec0, dtype=torch.uint8, pin_memory=True)
    host.copy_(d0)
    torch.cuda.synchronize()
    a, b = str(tmp_path / "a.meryl"), str(tmp_path / "b.meryl")
    with mb.Counter(21, 14) as c:
        c.add_bases_ptr(host.data_ptr(), n_reads * rec0, True)        # pinned: direct
        c.count()
        sa = c.stats()
        c.write_db(a)
    pageable = host.numpy().tobytes()
    with mb.Counter(21, 14) as c:
        c.add_bases(pageable, True)                                   # pageable: staged
        c.count()
        sb = c.stats()
        c.write_db(b)
    assert sa.nKmers == sb.nKmers == n_reads * (L - 21 + 1)
    util.assert_same_db(a, b)


def test_output_buffer_images_match_internal_images(tmp_path):
    """mb_set_output_buffer: data images land in the caller's buffer, byte-identical to the internal ones."""
    import torch
    data = synth.reads_fastq_bytes(100_000, 20_000, 150, 0.002, 13)
    with mb.Counter(21, 12) as c:
        c.add_file_bytes(data)
        c.count()
        want = [c.image(i) for i in range(129)]
    buf = torch.zeros(sum(len(x) for x in want[:64]) + 1024, dtype=torch.uint8, pin_memory=True)
    with mb.Counter(21, 12) as c:
        c.set_output_buffer(buf.data_ptr(), buf.numel())
        c.add_file_bytes(data)
        c.count()
        got = [c.image(i) for i in range(129)]
        out = str(tmp_path / "o.meryl")
        c.write_db(out)
    assert got == want
    raw = buf.numpy().tobytes()
    assert raw[: len(want[0])] == want[0]
    with mb.Counter(21, 12) as c:
        c.set_output_buffer(buf.data_ptr(), 1000)          # too small: loud error, no fallback
        c.add_file_bytes(data)
        with pytest.raises(mb.MerylB200Error) as e:
            c.count()
        assert e.value.code == -3


def test_suffix32_path_equals_full_key_path(tmp_path, monkeypatch):
    """k <= 21 on one GPU stores 32-bit suffix keys (half the sort traffic); MB_NO_SUFFIX32 forces the
    64-bit path.  Both must give the same bytes, also with several passes and several count groups."""
    data = synth.reads_fastq_bytes(400_000, 150_000, 150, 0.002, 91)
    dbs = []
    for env, ws in (("", 0), ("1", 0), ("", 1 << 28), ("1", 1 << 28)):
        if env:
            monkeypatch.setenv("MB_NO_SUFFIX32", env)
        else:
            monkeypatch.delenv("MB_NO_SUFFIX32", raising=False)
        if ws:                                           # (the in-place sort holds twice the keys per pass: cap the passes explicitly)
            monkeypatch.setenv("MB_MAX_PASS_KEYS", "6000000")
        out = str(tmp_path / ("db%d" % len(dbs)))
        with mb.Counter(21, 13, workspace_bytes=ws) as c:
            c.add_file_bytes(data)
            c.count()
            st = c.stats()
            c.write_db(out)
        assert st.keyBytes == (8 if env else 4)
        dbs.append((out, st.nPasses))
    for out, _ in dbs[1:]:
        util.assert_same_db(dbs[0][0], out)
    assert dbs[2][1] > 1 and dbs[3][1] > 1


@pytest.mark.parametrize("k", [31, 51])
def test_multi_pass_full_keys(tmp_path, k):
    """Several prefix-range passes with 64- and 128-bit keys (the in-place count path) == oracle."""
    data = synth.reads_fastq_bytes(200_000, 150_000, 150, 0.003, 300 + k)
    p = tmp_path / "r.fq"
    p.write_bytes(data)
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=12", "count", str(p), "output", ora])
    out = str(tmp_path / "gpu.meryl")
    with mb.Counter(k, 12, workspace_bytes=1 << 28) as c:
        c.add_file(str(p))
        c.count()
        assert c.stats().nPasses > 1
        c.write_db(out)
    util.assert_same_db(ora, out)

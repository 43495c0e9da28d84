"""The product's CUDA sources, run on the CPU.

This container has no GPU, so the kernels' indexing / synchronisation logic is checked here with the
cusim host emulator (tests/cusim: every CUDA thread is a fiber, barriers and warp collectives are
rendezvous points, TMA bulk copies are memcpys).  tests/cusim/Makefile rewrites every .cu file of
meryl_b200/csrc and builds a library with the same C ABI; these tests drive it exactly like the GPU
parity tests drive libmeryl_b200.so and compare with the oracle / golden dbs byte for byte.

This is TEST INFRASTRUCTURE: it proves the kernel logic, not the speed, and it is no fallback -- the
product library still refuses to run without a CUDA device (tests/test_host_logic.py)."""
import os
import subprocess

import numpy as np
import pytest

import util
import meryl_b200 as mb
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


def _vs_oracle(tmp_path, files, k, w_prefix, op="count", workspace_bytes=0, expect_key_bytes=None, min_passes=1):
    paths = []
    for name, data in files.items():
        p = tmp_path / name
        p.write_bytes(data)
        paths.append(str(p))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=%d" % w_prefix, op, *paths, "output", ora])
    with mb.Counter(k, w_prefix, op, workspace_bytes=workspace_bytes) as c:
        for p in paths:
            c.add_file(p)
        c.count()
        st = c.stats()
        c.write_db(out)
    util.assert_same_db(ora, out)
    if expect_key_bytes is not None:
        assert st.keyBytes == expect_key_bytes
    assert st.nPasses >= min_passes
    return st


@pytest.mark.parametrize("name", ["edge_k21", "edge_k8_simple"])
def test_sim_golden(tmp_path, name):
    inp, gold, words, _ = util.load_golden(name, str(tmp_path))
    k = int([w for w in words if w.startswith("k=")][0][2:])
    mem = float([w for w in words if w.startswith("memory=")][0][7:])
    out = str(tmp_path / "sim.meryl")
    mb.count([inp], out, k, "count", memory_gb=mem)
    util.assert_same_db(gold, out)


@pytest.mark.parametrize("k,w_prefix,key_bytes", [(21, 10, 4), (19, 8, 4), (22, 8, 8), (31, 7, 8), (32, 6, 8), (33, 6, 16), (51, 8, 16), (64, 6, 16), (8, 6, 4)])
def test_sim_k_sweep(tmp_path, k, w_prefix, key_bytes):
    """u32 suffix keys (2k - binBits <= 32), u64 keys and 128-bit keys through the same kernels."""
    st = _vs_oracle(tmp_path, {"a.fa": synth.fasta_bytes(2, 6_000, 100 + k)}, k, w_prefix, expect_key_bytes=key_bytes)
    assert st.nBatches == 1 and st.msMerge == 0             # one batch: nothing is merged


@pytest.mark.parametrize("op", ["count-forward", "count-reverse"])
def test_sim_forward_reverse(tmp_path, op):
    _vs_oracle(tmp_path, {"a.fq": synth.reads_fastq_bytes(3_000, 150, 120, 0.01, 5)}, 21, 8, op=op)


def test_sim_reads_with_errors_and_ragged_tiles(tmp_path):
    """More than one sort tile per segment (full + ragged tiles, look-back between tiles)."""
    _vs_oracle(tmp_path, {"r.fq": synth.reads_fastq_bytes(20_000, 3_000, 150, 0.003, 9)}, 21, 6, expect_key_bytes=8)
    (tmp_path / "b").mkdir()
    _vs_oracle(tmp_path / "b", {"r.fq": synth.reads_fastq_bytes(20_000, 3_000, 150, 0.003, 9)}, 19, 6, expect_key_bytes=4)


def test_sim_low_complexity(tmp_path):
    rng = np.random.default_rng(5)
    unit = synth.random_bases(500, rng).tobytes()
    recs = [b">r%d\n" % i + unit[: 100 + (1 + (i * 7) % 40) * 9] + b"\n" for i in range(60)]
    recs.append(b">polyA\n" + b"A" * 20_000 + b"\n")
    recs.append(b">polyAC\n" + b"AC" * 9_000 + b"\nNNNN\n")
    recs.append(b">short\nACG\n>empty\n\n")
    _vs_oracle(tmp_path, {"rep.fa": b"".join(recs)}, 21, 8)


@pytest.mark.parametrize("k,w_prefix,op", [(21, 12, "count"), (21, 9, "count-forward"), (22, 8, "count-reverse"), (31, 9, "count"),
                                           (64, 8, "count"), (5, 6, "count"), (33, 7, "count-reverse")])
def test_sim_multi_pass_prefilter(tmp_path, monkeypatch, k, w_prefix, op):
    """Several file-aligned passes: the extraction kernel's bit-sliced pass-membership test (first three bases of
    either strand) must select exactly the k-mers the rolling path selects -- N runs, lowercase, short and empty
    records, reads on both strands included."""
    monkeypatch.setenv("MB_MAX_PASS_KEYS", "150000")
    reads = synth.reads_fastq_bytes(20_000, 4_000, 150, 0.004, 77 + k)
    odd = b">n\nACGTNNNNNacgtacgtacgtacgtnnACGTACGTACGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGTTTTTTTT\n>s\nACG\n>e\n\n"
    _vs_oracle(tmp_path, {"r.fq": reads, "odd.fa": odd + synth.fasta_bytes(2, 3_000, k)}, k, w_prefix, op=op, min_passes=3)


@pytest.mark.parametrize("k,w_prefix,op", [(21, 12, "count"), (31, 9, "count-forward"), (18, 10, "count-reverse")])
def test_sim_dense_extraction_rounds(tmp_path, monkeypatch, k, w_prefix, op):
    """The dense form of a filtered pass (extract_dense_kernel) when a tile holds more members than one round takes:
    two passes of about half the k-mers each, 8 members per thread and round -> two rounds per tile."""
    monkeypatch.setenv("MB_MAX_PASS_KEYS", "300000")
    monkeypatch.setenv("MB_DENSE_ITEMS", "1")
    _vs_oracle(tmp_path, {"r.fq": synth.reads_fastq_bytes(20_000, 4_000, 150, 0.004, 177 + k)}, k, w_prefix, op=op, min_passes=2)


@pytest.mark.parametrize("cap", [0, 150_000])
def test_sim_streamed_output_dir(tmp_path, monkeypatch, cap):
    """mb_set_output_dir: the data files go from the device to the directory while the count runs (csrc/filesink.cpp, one
    worker per file); write_db adds the block indexes and merylIndex.  Same 129 files as the oracle, one and several passes;
    the streamed images are refused by mb_get_image; another directory is refused by write_db."""
    if cap:
        monkeypatch.setenv("MB_MAX_PASS_KEYS", str(cap))
    fq = tmp_path / "r.fq"
    fq.write_bytes(synth.reads_fastq_bytes(20_000, 4_000, 150, 0.004, 31))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=21", "wprefix=12", "count", str(fq), "output", ora])
    with mb.Counter(21, 12) as c:
        c.set_output_dir(out)
        c.add_file(str(fq))
        c.count()
        assert (c.stats().nPasses > 1) == bool(cap)
        with pytest.raises(mb.MerylB200Error):
            c.image(0)
        with pytest.raises(mb.MerylB200Error):
            c.write_db(str(tmp_path / "elsewhere.meryl"))
        c.write_db(out)
        assert c.image(64) and c.image(128)
    util.assert_same_db(ora, out)


@pytest.mark.parametrize("k,w_prefix,op,cap", [(21, 12, "count", 0), (20, 12, "count", 0), (17, 10, "count-forward", 150_000), (21, 14, "count-reverse", 150_000)])
def test_sim_in_place_group_sort(tmp_path, monkeypatch, k, w_prefix, op, cap):
    """u32 suffix keys sorted in place, group of bins by group of bins, against a scratch buffer (engine.cu): many small groups
    (MB_GROUP_KEYS), an even and an odd number of digit passes (k = 21: 32 bits, six passes, the group ends where it lay;
    k = 20 / 17: five / four passes), one and several passes over the reads.  Same db as the oracle, and as the ping-pong layout."""
    monkeypatch.setenv("MB_GROUP_KEYS", "3000")
    if cap:
        monkeypatch.setenv("MB_MAX_PASS_KEYS", str(cap))
    files = {"r.fq": synth.reads_fastq_bytes(20_000, 4_000, 150, 0.004, 500 + k)}
    st = _vs_oracle(tmp_path, files, k, w_prefix, op=op, expect_key_bytes=4, min_passes=2 if cap else 1)
    monkeypatch.setenv("MB_SORT_PINGPONG", "1")
    (tmp_path / "pp").mkdir()
    st2 = _vs_oracle(tmp_path / "pp", files, k, w_prefix, op=op, expect_key_bytes=4)
    assert st.nDistinct == st2.nDistinct and st.nLaunches > st2.nLaunches           # one sort per group, not per pass


def test_sim_multi_pass_rolling(tmp_path, monkeypatch):
    """The same kind of passes through the rolling path (filter switched off)."""
    monkeypatch.setenv("MB_MAX_PASS_KEYS", "150000")
    monkeypatch.setenv("MB_NO_PREFILTER", "1")
    _vs_oracle(tmp_path, {"r.fq": synth.reads_fastq_bytes(20_000, 4_000, 150, 0.004, 5)}, 21, 12, min_passes=3)


def test_sim_multi_pass(tmp_path):
    """A tiny arena forces several prefix-range passes (both key forms); the db must not change."""
    data = {"r.fq": synth.reads_fastq_bytes(30_000, 4_000, 150, 0.002, 77)}
    _vs_oracle(tmp_path, data, 21, 12, workspace_bytes=300 << 20, expect_key_bytes=4)


def test_sim_compress(tmp_path):
    """MB_FLAG_COMPRESS: homopolymer compression of file inputs (merylInput.C:258-262) on the host side of the ingest."""
    rng = np.random.default_rng(4)
    seq = bytearray()
    for _ in range(6000):
        seq += bytes([b"ACGTacgtN"[int(rng.integers(0, 9))]]) * int(rng.integers(1, 5))
    fa = tmp_path / "hp.fa"
    fa.write_bytes(b">a\n" + bytes(seq[:9000]) + b"\n" + bytes(seq[9000:]) + b"\n>b\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAACCCCCCCCCCCCCGT\n")
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=15", "wprefix=8", "compress", "count", str(fa), "output", ora])
    with mb.Counter(15, 8, flags=mb.FLAG_COMPRESS) as c:
        c.add_file(str(fa))
        c.count()
        c.write_db(out)
    util.assert_same_db(ora, out)


def test_sim_parallel_parse_in_the_pipeline(tmp_path, monkeypatch):
    """mb_add_file maps the file and parses it in blocks on several threads (cuts verified against the state machine)."""
    monkeypatch.setenv("MB_PARSE_THREADS", "4")
    monkeypatch.setenv("MB_PARSE_MIN_PIECE", "256")
    _vs_oracle(tmp_path, {"r.fq": synth.reads_fastq_bytes(5_000, 600, 150, 0.01, 3), "g.fa": synth.fasta_bytes(7, 2_000, 4, width=60, eol=b"\r\n")}, 21, 8)


# ---- count batches (W6: merylBlockWriter::finishBatch / mergeBatches, kmers-writer-block.C:172-386) --------------------

@pytest.mark.parametrize("k,w_prefix,op,batch_bases", [(21, 10, "count", 190_000), (51, 8, "count-forward", 150_000), (8, 6, "count-reverse", 100_000),
                                                       (3, 6, "count", 100_000)])          # (k = 3: prefix 6 is the whole k-mer)
def test_sim_count_batches(tmp_path, monkeypatch, k, w_prefix, op, batch_bases):
    """Reads that outgrow the device are counted in batches cut at record breaks; the batches' blocks are decoded, merged
    (saturating sums) and encoded again with the block writer's rules.  MB_MAX_BATCH_BASES makes a small input take that
    road; the db must be the single-batch db of the oracle, empty blocks and all."""
    monkeypatch.setenv("MB_MAX_BATCH_BASES", str(batch_bases))
    reads = synth.reads_fastq_bytes(20_000, 2_000, 150, 0.004, 7 + k)
    st = _vs_oracle(tmp_path, {"r.fq": reads, "g.fa": synth.fasta_bytes(2, 30_000, k)}, k, w_prefix, op=op)
    assert st.nBatches >= 2 and st.msMerge >= 0
    assert st.nKmers == 2_000 * (150 - k + 1) + 2 * (30_000 - k + 1)
    assert st.nTotal == st.nKmers


def test_sim_count_batches_pushed_bases_and_output_buffer(tmp_path, monkeypatch):
    """The same through mb_add_bases (one long push: the cut falls inside the call) with a caller-owned output buffer."""
    monkeypatch.setenv("MB_MAX_BATCH_BASES", "100000")
    monkeypatch.setenv("MB_RING_BYTES", "8192")               # the batches' images pass through small staging chunks
    rng = np.random.default_rng(11)
    recs = [synth.random_bases(int(rng.integers(30, 400)), rng).tobytes() for _ in range(1_600)]
    fa = tmp_path / "r.fa"
    fa.write_bytes(b"".join(b">r%d\n%s\n" % (i, r) for i, r in enumerate(recs)))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=25", "wprefix=8", "count", str(fa), "output", ora])
    stream = b".".join(recs)                                   # work-buffer form; the last record ends with the call
    buf = np.zeros(32 << 20, dtype=np.uint8)
    with mb.Counter(25, 8) as c:
        c.set_output_buffer(buf.ctypes.data, buf.size)
        c.add_bases(stream[:100_000], end_of_sequence=False)
        c.add_bases(stream[100_000:], end_of_sequence=True)
        c.count()
        st = c.stats()
        c.write_db(out)
    util.assert_same_db(ora, out)
    assert st.nBatches >= 3


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sim_file_shards_are_the_file(tmp_path, n_shards):
    """mb_add_file_shard (multi-GPU ingest: rank r takes shard r of every input): the shards, added one after the other to
    one context, are exactly the file -- 4-line FASTQ with quality lines that start with '@' and '+', multi-line FASTA,
    CRLF, a file shorter than the number of shards."""
    rng = np.random.default_rng(n_shards)
    fq = bytearray()
    for i in range(900):
        seq = synth.random_bases(int(rng.integers(40, 160)), rng).tobytes()
        q = bytes([b"@+I>"[int(rng.integers(0, 4))]]) + b"I" * (len(seq) - 1)
        fq += b"@r%d\n%s\n+\n%s\n" % (i, seq, q)
    fa = bytearray()
    for i in range(40):
        seq = synth.random_bases(int(rng.integers(100, 3000)), rng).tobytes()
        fa += b">s%d x\r\n" % i + b"\r\n".join(seq[j:j + 70] for j in range(0, len(seq), 70)) + b"\r\n"
    # (records without bases: after such a header the grammar reads the next line as sequence, so a header that follows a
    #  header is not taken as a cut)
    files = {"a.fq": bytes(fq), "b.fa": bytes(fa), "tiny.fa": b">t\nACGTACGTACGTACGTACGTACGTAC\n",
             "empties.fa": b">s0\nTTTGTGACGCGCGCCCNTTATGAAGGAGT\n>s1\n>s2\nAGAACGTCCTAGGGACCATTAGCCATGG\n>s3\n\n>s4\n>s5\nGGATCCATTGACCAGTTTGACCATTGACG\n"}
    paths = []
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
        paths.append(str(tmp_path / name))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=21", "wprefix=8", "count", *paths, "output", ora])
    with mb.Counter(21, 8) as c:
        for s in range(n_shards):
            for p in paths:
                c.add_file_shard(p, s, n_shards)
        c.count()
        c.write_db(out)
    util.assert_same_db(ora, out)


def test_sim_file_shard_exact_or_error(tmp_path):
    """Compressed files cannot be split.  For files the cut heuristic was not made for (multi-line FASTQ, quality lines that
    look like record starts, FASTA records inside a FASTQ file) the shards are still exactly the file -- or a shard reports
    that it does not end at a record boundary; never a silently different db."""
    gz = tmp_path / "r.fa.gz"
    gz.write_bytes(b"\x1f\x8b")
    with mb.Counter(21, 8) as c:
        with pytest.raises(mb.MerylB200Error) as e:
            c.add_file_shard(str(gz), 0, 2)
        assert "plain files" in str(e.value)
    rng = np.random.default_rng(9)
    errors = exact = 0
    for case in range(6):
        lines = [b"@r0", b"ACGTACGTACGTACGTACGTACGTACGTACGT", b"+", b"IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII"]     # (the format sniff wants one good record)
        for i in range(400):
            kind = int(rng.integers(0, 6))
            seq = synth.random_bases(int(rng.integers(30, 90)), rng).tobytes()
            lines.append([b"@r%d" % i, seq, b"+", b"@" + seq[1:], b"+" + seq[1:], b">x%d" % i][kind])
        f = tmp_path / ("soup%d.fq" % case)
        f.write_bytes(b"\n".join(lines) + b"\n")
        ora, out = str(tmp_path / ("ora%d.meryl" % case)), str(tmp_path / ("sim%d.meryl" % case))
        util.run_oracle(["k=21", "wprefix=8", "count", str(f), "output", ora])
        with mb.Counter(21, 8) as c:
            try:
                for s in range(5):
                    c.add_file_shard(str(f), s, 5)
            except mb.MerylB200Error as ex:
                assert "record boundary" in str(ex)
                errors += 1
                continue
            c.count()
            c.write_db(out)
        util.assert_same_db(ora, out)
        exact += 1
    assert errors + exact == 6


def test_sim_count_batches_edges(tmp_path, monkeypatch):
    """Batch cuts wait for a record break: a record longer than the batch limit stays whole; homopolymer compression
    (MB_FLAG_COMPRESS) carries no state over a cut; N runs and lowercase as usual."""
    monkeypatch.setenv("MB_MAX_BATCH_BASES", "30000")
    rng = np.random.default_rng(3)
    long_rec = synth.random_bases(100_000, rng).tobytes()                      # > 3 batch limits, no break inside
    recs = [b">long\n" + long_rec + b"\n"]
    for i in range(300):
        s = bytearray(synth.random_bases(int(rng.integers(50, 300)), rng).tobytes())
        if i % 7 == 0:
            s[10:20] = b"N" * 10
        if i % 5 == 0:
            s = bytearray(bytes(s).lower())
        recs.append(b">r%d\n" % i + bytes(s) + b"\n")
    fa = tmp_path / "e.fa"
    fa.write_bytes(b"".join(recs))
    for flags, words in ((0, []), (mb.FLAG_COMPRESS, ["compress"])):
        ora, out = str(tmp_path / ("ora%d.meryl" % flags)), str(tmp_path / ("sim%d.meryl" % flags))
        util.run_oracle(["k=19", "wprefix=8"] + words + ["count", str(fa), "output", ora])
        with mb.Counter(19, 8, flags=flags) as c:
            c.add_file(str(fa))
            c.count()
            st = c.stats()
            c.write_db(out)
        util.assert_same_db(ora, out)
        assert st.nBatches >= 3


def test_sim_compressed_input_with_a_quote_in_its_name(tmp_path):
    """A .gz input arrives through `gzip -dc '<path>'`: the path is shell-quoted, a quote inside it included."""
    import gzip
    data = synth.reads_fastq_bytes(4_000, 300, 150, 0.01, 5)
    plain, gz = tmp_path / "r.fq", tmp_path / "it's.fq.gz"
    plain.write_bytes(data)
    with gzip.open(gz, "wb") as fh:
        fh.write(data)
    outs = []
    for p in (plain, gz):
        out = str(tmp_path / (p.name + ".meryl"))
        with mb.Counter(21, 8, "count") as c:
            c.add_file(str(p))
            c.count()
            c.write_db(out)
        outs.append(out)
    util.assert_same_db(outs[0], outs[1])


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/meryl not built")
def test_sim_cli_compress_holds_for_the_inputs_that_follow(tmp_path):
    """`meryl k=K count a.fa compress b.fa output d`: the reference compresses homopolymer runs of the inputs named AFTER the
    word only (merylCommandBuilder.C:237-240, :546).  The command line tool (emulated library in front of it) writes the
    reference binary's directory, for the word in front of, between and behind the inputs."""
    rng = np.random.default_rng(21)

    def runs(n, seed_letters=b"ACGT"):
        out = bytearray()
        for _ in range(n):
            out += bytes([seed_letters[int(rng.integers(0, 4))]]) * int(rng.integers(1, 5))
        return bytes(out)
    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    a.write_bytes(b">a\n" + runs(4000) + b"\n>a2\n" + runs(500) + b"\n")
    b.write_bytes(b">b\n" + runs(3000) + b"\n")
    env = dict(os.environ, LD_PRELOAD=SIM_LIB)
    for i, words in enumerate((["compress", str(a), str(b)], [str(a), "compress", str(b)], [str(a), str(b), "compress"])):
        ref, out = str(tmp_path / ("ref%d.meryl" % i)), str(tmp_path / ("cli%d.meryl" % i))
        util.run_ref(["k=15", "memory=2", "threads=4", "count"] + words + ["output", ref])
        p = subprocess.run([mb.CLI_PATH, "k=15", "memory=2", "count"] + words + ["output", out], capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        util.assert_same_db(ref, out)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/meryl not built")
def test_sim_cli_count_suffix_in_front_of_k(tmp_path):
    """`count-suffix=GCG` named BEFORE `k=`: the reference builds the suffix while its k-mer masks are still zero
    (merylOp.H:139-147, kmers-tiny.H:86), only the last base survives and it keeps the k-mers that end with AAG.  The command
    line tool writes the same directory (and says so); behind `k=` the word means what it says.  Found by
    tools/fuzz_emulator.py --what cli."""
    fa = tmp_path / "a.fa"
    fa.write_bytes(synth.fasta_bytes(1, 20_000, 5))
    env = dict(os.environ, LD_PRELOAD=SIM_LIB)
    kept = []
    for i, words in enumerate((["count-suffix=GCG", "k=10"], ["k=10", "count-suffix=GCG"], ["count-suffix=gct", "k=10"], ["count-suffix=T", "k=10"])):
        ref, out = str(tmp_path / ("ref%d.meryl" % i)), str(tmp_path / ("cli%d.meryl" % i))
        util.run_ref(words + ["memory=2", "threads=4", "count-forward", str(fa), "output", ref])
        p = subprocess.run([mb.CLI_PATH] + words + ["memory=2", "count-forward", str(fa), "output", out], capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        util.assert_same_db(ref, out)
        kept.append(("NOTE: count-suffix= in front of k=" in p.stderr, {l.split()[0][-3:] for l in util.run_ref_stdout(["print", out]).splitlines()}))
    assert kept[0] == (True, {"AAG"}) and kept[1] == (False, {"GCG"}) and kept[2] == (True, {"AAT"}) and kept[3][0] is False

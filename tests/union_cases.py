"""Shared bodies of the union-sum parity tests: the same cases run against the emulated library on the CPU
(tests/test_cusim_union.py) and against the CUDA library on the GPU (tests/test_gpu_union.py)."""
import os

import pytest

import util
import meryl_b200 as mb
from meryl_b200 import synth


def oracle_db(tmp, name, data, k, op="count", wprefix=None):
    """A count db written by the C oracle (pinned to the reference by tests/test_oracle_vs_ref.py)."""
    p = os.path.join(tmp, name + ".fa")
    with open(p, "wb") as fh:
        fh.write(data)
    db = os.path.join(tmp, name + ".meryl")
    words = ["k=%d" % k, "memory=2"] + (["wprefix=%d" % wprefix] if wprefix else [])
    util.run_oracle(words + [op, p, "output", db])
    return db


def check_union(tmp, inputs, tag="u", op="union-sum", **kw):
    ora, out = os.path.join(tmp, tag + "_ora.meryl"), os.path.join(tmp, tag + "_dev.meryl")
    util.run_oracle([op, *inputs, "output", ora])
    st = mb.union_sum(inputs, out, **kw) if op == "union-sum" else mb.set_operation(op, inputs, out, **kw)
    util.assert_same_db(ora, out)
    return st, out


def case_golden(tmp, name):
    dbs, gold = util.load_union_golden(name, tmp)
    out = os.path.join(tmp, "dev.meryl")
    mb.union_sum(dbs, out)
    util.assert_same_db(gold, out)


def case_k_sweep(tmp, k, n=20_000):
    common = synth.fasta_bytes(1, n, 500 + k)
    a = oracle_db(tmp, "a", common + synth.fasta_bytes(1, n, 10 + k), k)
    b = oracle_db(tmp, "b", synth.fasta_bytes(1, n, 99 + k) + common + common, k, wprefix=8 if k < 12 else None)
    st, _ = check_union(tmp, [a, b])
    assert st.nOutputKmers < st.nInputKmers           # the common part merged


def case_union_of_unions(tmp, k=21, n=15_000):
    """Stream-written inputs (prefix 12, absent blocks), an input without k-mers, a single input, itself twice."""
    a = oracle_db(tmp, "a", synth.reads_fastq_bytes(n, n // 20, 150, 0.01, 1), k)
    b = oracle_db(tmp, "b", synth.reads_fastq_bytes(n, n // 20, 150, 0.01, 2), k, op="count-forward")
    e = oracle_db(tmp, "e", b">x\nACGT\n", k)
    _, u1 = check_union(tmp, [a, b], "u1")
    _, u2 = check_union(tmp, [e, b], "u2")
    for i, ins in enumerate([[u1, u2], [u1, a, e], [e, e], [u1], [a, a, a, a]]):
        check_union(tmp, ins, "c%d" % i)


def case_low_complexity(tmp, k=21):
    rep = b">r\n" + b"ACGTTGCA" * 3_000 + b"\n>p\n" + b"A" * 30_000 + b"\n"
    a = oracle_db(tmp, "a", rep + synth.fasta_bytes(1, 20_000, 3), k)
    b = oracle_db(tmp, "b", rep + rep, k)
    check_union(tmp, [a, b])


def case_groups(tmp, monkeypatch, k=31):
    """Several groups of data files (the device-memory plan), forced small."""
    a = oracle_db(tmp, "a", synth.fasta_bytes(2, 20_000, 7), k)
    b = oracle_db(tmp, "b", synth.fasta_bytes(2, 20_000, 8) + synth.fasta_bytes(1, 20_000, 7), k)
    monkeypatch.setenv("MB_UNION_MAX_FILES", "5")
    st, _ = check_union(tmp, [a, b])
    assert st.nGroups == 13


def case_set_operations(tmp, op, k=21, n=15_000):
    """The merge family (merylOp-nextMer.C:559-612) on one, two and three inputs that share k-mers with different counts,
    a 128-bit k, and an empty input."""
    common = synth.reads_fastq_bytes(n, n // 12, 150, 0.0, 1)
    a = oracle_db(tmp, "a", common + synth.reads_fastq_bytes(n, n // 25, 150, 0.0, 2), k)
    b = oracle_db(tmp, "b", common[: len(common) // 2] + synth.reads_fastq_bytes(n, n // 20, 150, 0.0, 3), k, op="count-forward")
    c = oracle_db(tmp, "c", common + common + synth.reads_fastq_bytes(n, n // 50, 150, 0.0, 2), k)
    e = oracle_db(tmp, "e", b">x\nACGT\n", k)
    cases = [[a, b], [c, a], [a], [a, e], [e, a], [a, b, c], [c, b, a, a]]     # (symmetric-difference of > 2: a composition of folds)
    for i, ins in enumerate(cases):
        check_union(tmp, ins, "s%d" % i, op=op)
    a2 = oracle_db(tmp, "a2", synth.fasta_bytes(2, n // 2, 5), 40)
    b2 = oracle_db(tmp, "b2", synth.fasta_bytes(1, n // 2, 5) + synth.fasta_bytes(1, n // 2, 6), 40)
    check_union(tmp, [a2, b2], "w", op=op)


VALUE_CASES = [("less-than", 3), ("greater-than", 1), ("at-least", 2), ("at-most", 1), ("equal-to", 2), ("not-equal-to", 2),
               ("increase", 7), ("increase", 4294967295), ("decrease", 1), ("multiply", 3), ("multiply", 4294967296),
               ("divide", 2), ("divide", 0), ("divide-round", 2), ("divide-round", 3), ("modulo", 2)]


def case_value_operations(tmp, n=15_000):
    """The one-input filters and arithmetic (merylOp-nextMer.C:497-557) on a db whose values range over 1..4."""
    common = synth.reads_fastq_bytes(n, n // 12, 150, 0.0, 1)
    a = oracle_db(tmp, "a", common + common + common[: len(common) // 2] + synth.reads_fastq_bytes(n, n // 25, 150, 0.0, 2), 21)
    for i, (op, param) in enumerate(VALUE_CASES):
        ora, out = os.path.join(tmp, "v%d_ora.meryl" % i), os.path.join(tmp, "v%d_dev.meryl" % i)
        util.run_oracle([op, str(param), a, "output", ora])
        mb.value_operation(op, param, a, out)
        util.assert_same_db(ora, out)


def case_small_ring(tmp, monkeypatch, k=21):
    """The data files' bytes travel through page-locked staging chunks in both directions (MB_RING_BYTES makes them
    small): chunks that hold several files, files that span several chunks, chunk edges inside a block, more chunks than
    ring buffers."""
    monkeypatch.setenv("MB_RING_BYTES", "4096")
    a = oracle_db(tmp, "a", synth.reads_fastq_bytes(30_000, 1_500, 150, 0.003, 5), k)
    b = oracle_db(tmp, "b", synth.reads_fastq_bytes(30_000, 1_000, 150, 0.003, 6), k)
    check_union(tmp, [a, b], tag="ring")
    monkeypatch.setenv("MB_UNION_MAX_FILES", "5")
    check_union(tmp, [a, b, a], tag="ring3")


def case_long_blocks(tmp, monkeypatch, k, decoder):
    """Few, long blocks (count dbs of prefix width 6 and 8: thousands of k-mers per block, several sub-blocks each, fields
    that hop to the next sub-block) through both block decoders: one thread per block, and one warp per block walking
    the unary chain in lock step over a shared-memory window (MB_DECODE picks one; the default goes by the block count)."""
    monkeypatch.setenv("MB_DECODE", decoder)
    a = oracle_db(tmp, "a", synth.reads_fastq_bytes(40_000, 600, 150, 0.004, 50 + k), k, wprefix=6)
    b = oracle_db(tmp, "b", synth.fasta_bytes(2, 30_000, k) + b">p\n" + b"ACGGT" * 2_000 + b"\n", k, wprefix=8)
    st, out = check_union(tmp, [a, b], tag="long")
    assert st.nInputKmers > 50_000
    check_union(tmp, [out, a], tag="long2")                  # and the stream writer's own blocks (prefix 12) as input


def case_threshold_words(tmp, n=12_000):
    """`distinct=` / `word-frequency=` thresholds of the filters (merylOperation::initializeThreshold): the number our
    mb_db_threshold derives from the input's histogram, fed to the filter, must give the reference binary's db."""
    import subprocess
    a = oracle_db(tmp, "a", synth.reads_fastq_bytes(n // 4, n // 40, 150, 0.01, 77), 21)
    for word, kw in (("distinct=0.5", dict(distinct=0.5)), ("distinct=0.98", dict(distinct=0.98)), ("d=0.0", dict(distinct=0.0)),
                     ("word-frequency=0.00002", dict(word_frequency=0.00002)), ("f=0.0000001", dict(word_frequency=0.0000001))):
        for op in ("greater-than", "at-most"):
            t = mb.db_threshold(a, **kw)
            ref, out = os.path.join(tmp, "ref_%s_%s.meryl" % (op, word)), os.path.join(tmp, "dev_%s_%s.meryl" % (op, word))
            p = subprocess.run([util.REF_BIN, op, word, a, "output", ref], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-500:]
            mb.value_operation(op, t, a, out)
            util.assert_same_db(ref, out)


def case_count_suffix(tmp, n=8_000, full=False):
    """`count-suffix=`: the reference counts only the k-mers that end with the given bases (simple mode, its own db geometry);
    here: the full count, then mb_keep_suffix.  Same directory as the reference binary's, for canonical / forward / reverse
    counts, prefix widths 6 and above, and a suffix no k-mer has."""
    import subprocess
    fq = os.path.join(tmp, "r.fq")
    with open(fq, "wb") as fh:
        fh.write(synth.reads_fastq_bytes(n, n // 20, 150, 0.004, 31))
    # (the reference's simple counter allocates 2 * 4^(k-s) bytes: keep k - s small)
    cases = [(12, "ACG", "count"), (13, "tg", "count-forward"), (12, "A", "count-reverse"), (9, "CCCCCC", "count")]
    if full:
        cases += [(21, "GATTACA", "count"), (15, "TG", "count")]
    for k, suffix, op in cases:
        tag = "%d_%s" % (k, suffix)
        ref, full_db, out = (os.path.join(tmp, x + tag + ".meryl") for x in ("ref", "full", "dev"))
        p = subprocess.run([util.REF_BIN, "k=%d" % k, "memory=2", "count-suffix=" + suffix, op, fq, "output", ref], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-500:]
        with mb.Counter(k, 8, op) as c:                      # (the full count's own prefix width does not matter)
            c.add_file(fq)
            c.count()
            c.write_db(full_db)
        mb.keep_suffix(full_db, suffix, out)
        util.assert_same_db(ref, out)


_CODE = {"A": 0, "C": 1, "T": 2, "G": 3}


def _pack(kmer: str):
    v = 0
    for ch in kmer:
        v = (v << 2) | _CODE[ch]
    return v & 0xffffffffffffffff, v >> 64


def case_lookup(tmp, k, n=6_000, monkeypatch=None):
    """mb_lookup (merylExactLookup::value): every k-mer the reference binary prints for a database must come back with its
    value, neighbours of present k-mers and random k-mers that are absent with 0, values outside [min, max] with 0 -- count
    dbs (block writer, several prefix widths) and a union's output (stream writer, prefix 12); with the device arena cut down
    so that the files are decoded in several groups."""
    import subprocess
    import numpy as np
    a = oracle_db(tmp, "a", synth.reads_fastq_bytes(n, n // 15, 150, 0.004, 40 + k), k)
    b = oracle_db(tmp, "b", synth.fasta_bytes(2, n, k), k, wprefix=7)
    _, u = check_union(tmp, [a, b], tag="lk")
    rng = np.random.default_rng(k)
    for db, ws in ((a, 0), (b, 0), (u, 0), (a, 40 << 20)):
        if ws and monkeypatch is not None:
            monkeypatch.setenv("MB_UNION_MAX_FILES", "7")      # ... in ten groups of files
        p = subprocess.run([util.REF_BIN, "print", db], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-300:]
        truth = {}
        for line in p.stdout.split("\n"):
            if line:
                kmer, val = line.split()
                truth[_pack(kmer)] = int(val)
        assert truth
        present = list(truth)
        queries = present + [(lo ^ 1, hi) for lo, hi in present[:2000]] + \
            [(int(rng.integers(0, 2 ** 63)) & ((1 << min(64, 2 * k)) - 1), int(rng.integers(0, 2 ** 62)) & ((1 << max(0, 2 * k - 64)) - 1)) for _ in range(2000)]
        lo = np.array([q[0] for q in queries], dtype=np.uint64)
        hi = np.array([q[1] for q in queries], dtype=np.uint64) if k > 32 else None
        got = mb.lookup(db, lo, hi, workspace_bytes=ws)
        want = np.array([truth.get(q, 0) for q in queries], dtype=np.uint32)
        assert np.array_equal(got, want)
        # the whole database in order (mb_for_each_kmer, the walk behind `meryl print`) is the reference's print, sorted
        alo, ahi, aval = mb.all_kmers(db, workspace_bytes=ws)
        listed = [((int(alo[i]), int(ahi[i]) if ahi is not None else 0), int(aval[i])) for i in range(alo.size)]
        assert listed == sorted(truth.items(), key=lambda kv: (kv[0][1], kv[0][0]))
        got = mb.lookup(db, lo, hi, min_value=2, max_value=3, workspace_bytes=ws)
        want = np.array([truth.get(q, 0) if 2 <= truth.get(q, 0) <= 3 else 0 for q in queries], dtype=np.uint32)
        assert np.array_equal(got, want)
        if not ws:
            with mb.Table(db, min_value=2, max_value=3) as t:                # the same from the resident table
                assert np.array_equal(t.values(lo, hi), want)


def case_cli_print(tmp, k, n=3_000, env=None):
    """`meryl-b200 print db`: blocks decoded on the device, text made on the host; the reference's lines (its threads
    interleave the files, so the reference side is compared sorted; ours is ascending in meryl's code A < C < T < G)."""
    import subprocess
    db = oracle_db(tmp, "a", synth.reads_fastq_bytes(n, max(200, n // 15), 150, 0.004, k), k)
    ours = subprocess.run([mb.CLI_PATH, "print", db], capture_output=True, text=True, env=env)
    ref = subprocess.run([util.REF_BIN, "print", db], capture_output=True, text=True)
    assert ours.returncode == 0 and ref.returncode == 0, ours.stderr[-500:]
    mine = ours.stdout.splitlines()
    code = str.maketrans("ACTG", "0123")
    assert mine and sorted(mine, key=lambda l: l.split("\t")[0].translate(code)) == mine
    assert sorted(mine) == sorted(ref.stdout.splitlines())


def _existence_queries(k, n, seed):
    """Sequences for meryl-lookup: reads of the database's genome (found), of another genome (not found), lower case, Ns,
    shorter than k, exactly k, empty; names with flags after a blank or a tab."""
    import numpy as np
    rng = np.random.default_rng(seed)
    seqs = []
    fq = synth.reads_fastq_bytes(n, 60, 150, 0.01, 40 + k).split(b"\n")            # same genome seed as the database below
    seqs += [fq[i] for i in range(1, len(fq) - 1, 4)]
    other = synth.reads_fastq_bytes(n, 20, 150, 0.0, 977).split(b"\n")
    seqs += [other[i] for i in range(1, len(other) - 1, 4)]
    seqs.append(seqs[0].lower())
    seqs.append(seqs[1][:70] + b"N" + seqs[1][70:100] + b"nn" + seqs[1][100:])
    seqs.append(seqs[2][:k - 1])
    seqs.append(seqs[3][:k])
    seqs.append(b"")
    seqs.append(seqs[4][:k + 5] + b"R" + seqs[5][:k] + b"-" + seqs[6])
    long_one = synth.fasta_bytes(1, 3_000, 40 + k).split(b"\n", 1)[1].replace(b"\n", b"")
    seqs.append(long_one + seqs[7] + long_one[::-1])
    seqs.append(b"")
    names = []
    for i in range(len(seqs)):
        names.append(b"q%d" % i + (b"" if i % 3 == 0 else b" len=%d x" % len(seqs[i]) if i % 3 == 1 else b"\tflag"))
    return names, seqs


def _ref_existence(tmp, fa, dbs, extra=()):
    import subprocess
    out = os.path.join(tmp, "ref_existence.txt")
    p = subprocess.run([util.REF_LOOKUP_BIN, "-existence", "-sequence", fa, "-mers", *dbs, "-output", out, "-threads", "4", *extra],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-500:]
    with open(out) as fh:
        return [line.rstrip("\n").split("\t") for line in fh]


def case_existence(tmp, k, n=6_000, monkeypatch=None):
    """mb_existence (meryl-lookup -existence, existence.C:59-86): per sequence the number of k-mers and the number found in the
    database in either orientation, against the reference's meryl-lookup binary -- a canonical count db, a forward count db
    (only one strand's form is stored: the reverse-complement lookup matters), a union's output; a value window (-min/-max);
    the files decoded in several groups; an empty database."""
    import numpy as np
    a = oracle_db(tmp, "a", synth.reads_fastq_bytes(n, n // 15, 150, 0.004, 40 + k), k)
    f = oracle_db(tmp, "f", synth.reads_fastq_bytes(n, n // 30, 150, 0.004, 40 + k), k, op="count-forward")
    e = oracle_db(tmp, "e", b">none\nACGT\n", k)
    names, seqs = _existence_queries(k, n, k)
    fa = os.path.join(tmp, "queries.fa")
    with open(fa, "wb") as fh:
        for i, (nm, sq) in enumerate(zip(names, seqs)):
            if i % 2 == 0:
                fh.write(b">" + nm + b"\n" + b"\n".join(sq[j:j + 60] for j in range(0, len(sq), 60)) + b"\n")
            else:
                fh.write(b">" + nm + b"\r\n" + sq + b"\r\n")
    for dbs, ws, extra, window in (([a, f], 0, (), (0, 0xffffffff)), ([a], 40 << 20, ("-min", "2", "-max", "3"), (2, 3)), ([e], 0, (), (0, 0xffffffff))):
        if ws and monkeypatch is not None:
            monkeypatch.setenv("MB_UNION_MAX_FILES", "7")
        ref = _ref_existence(tmp, fa, dbs, extra)
        assert [r[0].encode() for r in ref] == [nm.split()[0] for nm in names]
        for d, db in enumerate(dbs):
            tot, fnd, ndb = mb.existence(db, seqs, min_value=window[0], max_value=window[1], workspace_bytes=ws)
            assert [int(x) for x in tot] == [int(r[1]) for r in ref]
            assert [int(x) for x in fnd] == [int(r[3 + 2 * d]) for r in ref]
            assert ndb == int(ref[0][2 + 2 * d])
        if dbs == [a, f]:
            assert sum(int(r[3]) for r in ref) > 0 and sum(int(r[5]) for r in ref) > 0
    # -10x: the k-mers that start in the first 23 positions are not looked up
    tot0, fnd0, _ = mb.existence(a, seqs)
    tot1, fnd1, _ = mb.existence(a, seqs, skip_first=23)
    tot2, fnd2, _ = mb.existence(a, [s[23:] if len(s) >= 23 + k else b"" for s in seqs])
    assert np.array_equal(tot0, tot1) and np.array_equal(fnd1, fnd2) and int(fnd1.sum()) < int(fnd0.sum())
    # the per-position form agrees with the per-sequence sums; a found position has a non-zero value sum
    flags, vals = mb.position_lookup(a, seqs, workspace_bytes=40 << 20)
    edges = np.cumsum([0] + [len(s) for s in seqs])
    assert [int((flags[edges[i]:edges[i + 1]] & 1).sum()) for i in range(len(seqs))] == [int(x) for x in tot0]
    assert [int(((flags[edges[i]:edges[i + 1]] >> 1) & 1).sum()) for i in range(len(seqs))] == [int(x) for x in fnd0]
    assert np.array_equal(vals > 0, (flags & 2) > 0)
    # the resident table (mb_table_*): the same answers, asked in several batches against one decoded copy
    with mb.Table(a) as t:
        assert t.k == k and t.kmers == mb.existence(a, [])[2]
        half = len(seqs) // 2
        t1, f1, fl1, v1 = t.query(seqs[:half], values=True)
        t2, f2, fl2, v2 = t.query(seqs[half:], values=True)
        assert np.array_equal(np.concatenate([t1, t2]), tot0) and np.array_equal(np.concatenate([f1, f2]), fnd0)
        assert np.array_equal(np.concatenate([fl1, fl2]), flags) and np.array_equal(np.concatenate([v1, v2]), vals)
        assert np.array_equal(t.query(seqs, skip_first=23)[1], fnd1)
        assert t.query([])[0].size == 0
    with mb.Table(a, min_value=2, max_value=3) as t:
        w_tot, w_fnd, w_ndb = mb.existence(a, seqs, min_value=2, max_value=3)
        assert t.kmers == w_ndb and np.array_equal(t.query(seqs)[1], w_fnd)
    with mb.Table(e) as t:                                   # a database without k-mers
        assert t.kmers == 0 and np.array_equal(t.query(seqs)[0], tot0) and int(t.query(seqs)[1].sum()) == 0
    with pytest.raises(mb.MerylB200Error):                   # no room to keep it resident: the one-shot calls are the fallback
        mb.Table(a, workspace_bytes=1 << 20)
    # no sequences at all: only the database's size comes back
    tot, fnd, ndb = mb.existence(a, [])
    assert tot.size == 0 and ndb > 0


def case_cli_lookup(tmp, k, n=6_000, env=None):
    """`meryl-lookup-b200` against the reference's meryl-lookup, file for file: -existence on a FASTA file with two databases;
    -include / -exclude on FASTQ reads (single, paired, -10x, a value window) and on FASTA; -bed / -bed-runs (with and without
    labels) / -wig-count / -wig-depth; usage errors."""
    import subprocess
    a = oracle_db(tmp, "a", synth.reads_fastq_bytes(n, n // 15, 150, 0.004, 40 + k), k)
    f = oracle_db(tmp, "f", synth.reads_fastq_bytes(n, n // 30, 150, 0.004, 40 + k), k, op="count-forward")
    names, seqs = _existence_queries(k, n, k)
    fa = os.path.join(tmp, "queries.fa")
    with open(fa, "wb") as fh:
        for i, (nm, sq) in enumerate(zip(names, seqs)):
            if sq or i % 2:                                  # (the reference's FASTA writer divides by the length: no empty sequences here)
                fh.write(b">" + nm + b"\n" + b"\n".join(sq[j:j + 60] for j in range(0, len(sq), 60)) + b"\n")
    fa_all = os.path.join(tmp, "queries_all.fa")
    with open(fa_all, "wb") as fh:
        for nm, sq in zip(names, seqs):
            fh.write(b">" + nm + b"\n" + sq + b"\n")
    # reads: half from the database's genome, half from another; as FASTQ (qualities starting above '!' and, for a few, with '!')
    r1, r2 = os.path.join(tmp, "r1.fq"), os.path.join(tmp, "r2.fq")
    own = synth.reads_fastq_bytes(n, 40, 100, 0.02, 40 + k).split(b"\n")
    oth = synth.reads_fastq_bytes(n, 40, 100, 0.0, 4242).split(b"\n")
    with open(r1, "wb") as f1, open(r2, "wb") as f2:
        for i in range(40):
            s1 = own[4 * i + 1] if i % 2 else oth[4 * i + 1]
            s2 = oth[4 * i + 1][::-1] if i % 3 else own[4 * i + 1][:60]
            q1 = (b"!" if i % 7 == 0 else b"I") + b"F" * (len(s1) - 1)
            f1.write(b"@pair%d/1 extra words\n" % i + s1 + b"\n+\n" + q1 + b"\n")
            f2.write(b"@pair%d/2\n" % i + s2 + b"\n+pair%d/2\n" % i + b"5" * len(s2) + b"\n")

    def both(words, outs):
        got = []
        for exe, tag in ((util.REF_LOOKUP_BIN, "ref"), (mb.LOOKUP_CLI_PATH, "dev")):
            names_ = [os.path.join(tmp, "%s_%s" % (tag, o)) for o in outs]
            # (-threads: the reference sizes its load buffers at 512 MB per thread and refuses when that exceeds the memory limit)
            p = subprocess.run([exe] + words + ["-threads", "4", "-output"] + names_, capture_output=True, text=True, env=env if tag == "dev" else None)
            assert p.returncode == 0, (tag, p.stderr[-500:])
            got.append([open(x, "rb").read() for x in names_])
        assert got[0] == got[1], words
        return got[1]

    out = both(["-existence", "-sequence", fa_all, "-mers", a, f], ["ex.txt"])
    assert out[0].count(b"\n") == len(seqs)
    both(["-existence", "-sequence", r1, "-mers", f, "-min", "2"], ["ex2.txt"])
    inc = both(["-include", "-sequence", r1, "-mers", a], ["inc.fq"])
    exc = both(["-exclude", "-sequence", r1, "-mers", a], ["exc.fq"])
    assert inc[0].count(b"nKmers=") + exc[0].count(b"nKmers=") == 40 and inc[0].startswith(b"@pair1/1 nKmers=") and b">pair" in inc[0] + exc[0]
    both(["-include", "-sequence", r1, r2, "-mers", a], ["inc1.fq", "inc2.fq"])
    both(["-exclude", "-10x", "-sequence", r1, r2, "-mers", a, "-min", "2", "-max", "50"], ["exc1.fq", "exc2.fq"])
    both(["-include", "-sequence", fa, "-mers", f], ["inc.fa"])
    # the per-position reports (dump.C), on sequences with hits, misses, Ns, lower case, and sequences shorter than k
    for rep in (["-bed"], ["-bed-runs"], ["-wig-count"], ["-wig-depth"], ["-bed", "-labels", "A", "B"], ["-bed-runs", "-labels", "onlyA"]):
        out = both(rep + ["-sequence", fa_all, "-mers", a, f], ["pos.txt"])
        assert out[0]
        both(rep + ["-sequence", r1, "-mers", f, a, "-min", "2", "-max", "40"], ["pos2.txt"])
    pr = subprocess.run([util.REF_LOOKUP_BIN, "-existence", "-sequence", fa, "-mers", a, "-threads", "4"], capture_output=True)   # no -output: stdout
    pd = subprocess.run([mb.LOOKUP_CLI_PATH, "-existence", "-sequence", fa, "-mers", a, "-threads", "4"], capture_output=True, env=env)
    assert pr.returncode == pd.returncode == 0 and pr.stdout == pd.stdout and pd.stdout
    pd = subprocess.run([mb.CLI_PATH, "lookup", "-existence", "-sequence", fa, "-mers", a], capture_output=True, env=env)  # the same through `meryl-b200 lookup`
    assert pd.returncode == 0 and pr.stdout == pd.stdout
    both(["-exclude", "-sequence", fa, "-mers", f, "-threads", "2", "-memory", "4"], ["exc.fa"])
    # the sequences in many small batches against the resident table, and the fallback without it (one-shot calls per batch)
    for knobs in ({"MB_LOOKUP_BATCH": "2000"}, {"MB_LOOKUP_BATCH": "5000", "MB_LOOKUP_NO_TABLE": "1"}):
        e2 = dict(env if env is not None else os.environ, **knobs)
        for rep, ref_name in ((["-existence"], "dev_ex.txt"), (["-wig-count"], None), (["-bed-runs", "-labels", "A", "B"], None)):
            o = os.path.join(tmp, "batched.txt")
            pd = subprocess.run([mb.LOOKUP_CLI_PATH] + rep + ["-sequence", fa_all, "-mers", a, f, "-output", o], capture_output=True, env=e2)
            assert pd.returncode == 0, pd.stderr[-300:]
            if ref_name is None:
                pr = subprocess.run([util.REF_LOOKUP_BIN] + rep + ["-sequence", fa_all, "-mers", a, f, "-threads", "4", "-output", o + ".ref"], capture_output=True)
                assert pr.returncode == 0
                ref_name = o + ".ref"
            assert open(o, "rb").read() == open(os.path.join(tmp, ref_name), "rb").read(), (knobs, rep)
    # a gzip-compressed input (read through `gzip -dc`), with a quote in its name; gzip-compressed output
    import gzip
    gz = os.path.join(tmp, "que'ries.fa.gz")
    with gzip.open(gz, "wb") as fh:
        fh.write(open(fa_all, "rb").read())
    outgz = os.path.join(tmp, "ex'out.txt.gz")
    pd = subprocess.run([mb.LOOKUP_CLI_PATH, "-existence", "-sequence", gz, "-mers", a, f, "-output", outgz], capture_output=True, env=env)
    assert pd.returncode == 0, pd.stderr[-300:]
    assert gzip.open(outgz, "rb").read() == open(os.path.join(tmp, "dev_ex.txt"), "rb").read()
    # usage errors (meryl-lookup.C:291-368): same sentences, exit code 1
    for words in (["-sequence", fa, "-mers", a, "-output", "x"], ["-existence", "-mers", a, "-output", "x"],
                  ["-existence", "-sequence", fa, "-output", "x"],
                  ["-include", "-sequence", fa, "-mers", a, f, "-output", "x"], ["-existence", "-sequence", fa, fa, "-mers", a, "-output", "x"],
                  ["-include", "-sequence", r1, r2, "-mers", a, "-output", "x"], ["-existence", "-frobnicate"],
                  ["-wig-count", "-sequence", fa, "-mers", a, "-labels", "A", "-output", "x"], ["-bed", "-sequence", fa, fa, "-mers", a, "-output", "x", "y"]):
        pr = subprocess.run([util.REF_LOOKUP_BIN] + words, capture_output=True, text=True, cwd=tmp)
        pd = subprocess.run([mb.LOOKUP_CLI_PATH] + words, capture_output=True, text=True, cwd=tmp, env=env)
        assert pr.returncode == pd.returncode == 1
        sentences = [l for l in pr.stderr.splitlines() if l.startswith(("No ", "Only ", "Unknown ", "Labels "))]
        assert sentences and all(l in pd.stderr for l in sentences), (words, pr.stderr[-300:], pd.stderr)


def case_stream_block_split(tmp, monkeypatch, k, limit):
    """A prefix with more k-mers than the stream writer keeps in a block (16 Mi in the reference,
    kmers-writer-stream.C:127-141) is written as consecutive blocks of the same prefix under one index entry.  The limit
    is lowered on both sides (MB_STREAM_BLOCK_MAX / ORA_STREAM_BATCH_MAX) so that most of the 4096 prefixes are cut in
    two to five parts, one of them into parts of exactly the limit; tests/test_gpu_zzzz_late.py holds the real limit
    against the reference binary.  Written (encoder) and read back (block tables from the container headers): as the input
    of another operation, by the database walk, by the exact lookup."""
    import numpy as np
    monkeypatch.setenv("MB_STREAM_BLOCK_MAX", str(limit))
    monkeypatch.setenv("ORA_STREAM_BATCH_MAX", str(limit))
    def index_of(db):
        return np.concatenate([np.fromfile(os.path.join(db, n), dtype=np.uint64) for n in sorted(os.listdir(db)) if n.endswith(".merylIndex")]).reshape(-1, 3)
    fa, fb = synth.fasta_bytes(2, 150_000, 40 + k), synth.fasta_bytes(1, 150_000, 41 + k) + synth.fasta_bytes(1, 150_000, 40 + k)
    b = oracle_db(tmp, "b", fb, k, op="count-forward")
    # one prefix (GATTAC = 0xca1) is filled up to a multiple of the limit: its last part is a full one
    _, u0 = check_union(tmp, [oracle_db(tmp, "a0", fa, k, op="count-forward"), b], "before")
    ix = index_of(u0)
    n0 = int(ix[ix[:, 0] == 0xca1, 2].sum())
    fill = (n0 // limit + 2) * limit - n0
    tails = np.random.default_rng(limit + k).choice(4 ** 9, size=fill, replace=False)
    extra = b"".join(b">e\n" + b"GATTAC" + bytes(b"ACGT"[(int(t) >> (2 * j)) & 3] for j in range(9)) + b"C" * (k - 15) + b"\n" for t in tails)
    a = oracle_db(tmp, "a", fa + extra, k, op="count-forward")
    st, u = check_union(tmp, [a, b], "split")
    assert st.nOutputKmers > 4096 * limit                  # the average prefix is over the limit
    ix = index_of(u)
    assert (ix[:, 2] > limit).sum() > 1000 and (ix[:, 2] > 3 * limit).any()
    assert int(ix[ix[:, 0] == 0xca1, 2][0]) == n0 + fill and (n0 + fill) % limit == 0
    _, u2 = check_union(tmp, [u, a], "again")              # split blocks as input (+ written again)
    check_union(tmp, [u2, u, b], "three", op="intersect-sum")
    lo, hi, val = mb.all_kmers(u)
    assert lo.size == st.nOutputKmers and int(val.sum(dtype=np.uint64)) == st.nTotal
    key = lo if hi is None else hi.astype(object) * (1 << 64) + lo.astype(object)
    assert all(key[i] < key[i + 1] for i in range(0, key.size - 1, max(1, key.size // 5000)))
    pick = np.arange(0, lo.size, max(1, lo.size // 3000))
    got = mb.lookup(u, lo[pick], None if hi is None else hi[pick])
    assert (got == val[pick]).all()


def case_stream_block_real_limit(tmp_path):
    """16 Mi + 300 000 distinct 19-mers that share their first six bases (one 12-bit prefix of a set operation's output) are
    written as a block of 16 Mi k-mers followed by a block of 300 000 under one index entry (merylStreamWriter::addMer,
    kmers-writer-stream.C:127-141; merylFileIndex::set, kmers-files.H:116-128) -- byte for byte what the reference binary
    writes -- and such a database is read back."""
    import numpy as np
    k, n = 19, (16 << 20) + 300_000
    rec = np.empty((n, 23), dtype=np.uint8)
    rec[:, :3] = np.frombuffer(b">s\n", dtype=np.uint8)
    rec[:, 3:9] = np.frombuffer(b"GATTAC", dtype=np.uint8)
    i = np.arange(n, dtype=np.uint32)
    for j in range(13):                                     # the tail spells the record's number: all k-mers distinct
        rec[:, 9 + j] = np.frombuffer(b"ACTG", dtype=np.uint8)[(i >> (2 * (12 - j))) & 3]      # meryl: A 0, C 1, T 2, G 3
    rec[:, 22] = ord("\n")
    fa = str(tmp_path / "one_prefix.fa")
    with open(fa, "wb") as fh:
        fh.write(b">other\nTTGACCATGGCATTACGGATCAGGCATTTAGC\n")
        rec.tofile(fh)
    del rec
    a, ref = str(tmp_path / "a.meryl"), str(tmp_path / "ref.meryl")
    util.run_ref(["k=%d" % k, "memory=8", "threads=4", "count-forward", fa, "output", a], timeout=1800)
    util.run_ref(["union-sum", a, "output", ref], timeout=1800)
    ix = np.fromfile(os.path.join(ref, "0x110010.merylIndex"), dtype=np.uint64).reshape(-1, 3)     # GATTAC = 110010 100001
    assert ix[0b100001, 2] == n                             # one index entry for the whole prefix

    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["union-sum", ref, "output", ora])      # the oracle's restatement of the limit, read and written
    util.assert_same_db(ref, ora)

    out = str(tmp_path / "dev.meryl")
    st = mb.union_sum([a], out)                             # written: two blocks of one prefix
    assert st.nOutputKmers == n + 14
    util.assert_same_db(ref, out)
    again = str(tmp_path / "dev2.meryl")
    mb.union_sum([ref], again)                              # read: block places from the container headers
    util.assert_same_db(ref, again)
    lo = np.array([0b110010100001 << 26, (0b110010100001 << 26) + n - 1, (0b110010100001 << 26) + n], dtype=np.uint64)
    assert mb.lookup(ref, lo).tolist() == [1, 1, 0]

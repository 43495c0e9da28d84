#!/usr/bin/env python
"""Randomised soak of the count path (with --what union: of chains of set / value operations; with --what multi: of the collective count of 2-8 emulated ranks; with --what cli / lookup: of the command line tools next to the reference's binaries) in the emulator (tests/cusim) against the C oracle: random k, operation, prefix width,
inputs with every oddity of the parser's grammar, and random settings of the testing knobs (passes, groups, batches, sweep
shape, dense extraction, key width).  Every case must give the oracle's 129 files byte for byte.  TEST INFRASTRUCTURE -- it
proves logic in a container without a GPU; `MB_TEST_LIB=... pytest -m gpu` and this script are the two checks used for
host-side changes made when no GPU minutes are left.

    python tools/fuzz_emulator.py --seconds 600 --jobs 6 [--seed 1]

A failing case is kept under --keep (inputs, both databases, the environment) and reported; exit code 1."""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SIM_DIR = os.path.join(ROOT, "tests", "cusim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libmeryl_b200_sim.so")
KNOBS = ("MB_MAX_PASS_KEYS", "MB_GROUP_KEYS", "MB_DENSE_ITEMS", "MB_SORT", "MB_MAX_BATCH_BASES", "MB_EXTRACT_DENSE", "MB_SORT_PINGPONG",
         "MB_COUNT_GROUP", "MB_HIST_SINGLE", "MB_NO_SUFFIX32", "MB_NO_PREFILTER", "MB_PARSE_MIN_PIECE", "MB_PARSE_THREADS")


def make_input(rng, k):
    """One sequence file: FASTA or FASTQ with the oddities bufSeqFile::loadBases accepts."""
    import numpy as np
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    eol = b"\r\n" if rng.random() < 0.15 else b"\n"
    fastq = rng.random() < 0.5
    out = []
    nrec = int(rng.integers(1, 40))
    scale = int(rng.choice([30, 200, 2000, 20000]))
    genome = acgt[rng.integers(0, 4, size=scale * 4 + 200, dtype=np.uint8)]
    for r in range(nrec):
        n = int(rng.integers(0, scale))
        if rng.random() < 0.1:
            n = int(rng.integers(max(0, k - 2), k + 3))                 # around k: zero, one, two k-mers
        at = int(rng.integers(0, genome.size - n)) if rng.random() < 0.7 else 0      # repeats between the records: counts > 1
        s = genome[at: at + n].copy()
        if rng.random() < 0.3 and n:
            s[rng.random(n) < 0.02] = ord("N")
        if rng.random() < 0.2 and n:
            low = rng.random(n) < 0.3
            s[low] = s[low] | 0x20
        if rng.random() < 0.1 and n > 40:
            p = int(rng.integers(0, n - 30))
            s[p: p + 30] = s[p]                                          # a homopolymer run
        s = s.tobytes()
        if fastq:
            out += [b"@r%d some text" % r + eol, s + eol, b"+" + (b"r%d" % r if rng.random() < 0.3 else b"") + eol,
                    ((b"@" if rng.random() < 0.2 else b"I") + b"I" * (n - 1) if n else b"") + eol]
        else:
            out.append(b">s%d" % r + eol)
            w = int(rng.choice([60, 80, 7, 100000]))
            for p in range(0, n, w):
                out.append(s[p: p + w] + eol)
            if rng.random() < 0.1:
                out.append(eol)
    if rng.random() < 0.1 and out:
        out[-1] = out[-1].rstrip(b"\r\n")                                # no end of line at the end of the file
    return b"".join(out), (".fq" if fastq else ".fa")


def one_case(seed, keep):
    import numpy as np
    import util
    import meryl_b200 as mb
    mb.LIB_PATH, mb._lib = SIM_LIB, None
    rng = np.random.default_rng(seed)
    k = int(rng.choice([3, 4, 8, 11, 15, 16, 17, 19, 20, 21, 22, 23, 27, 31, 32, 33, 40, 48, 51, 63, 64]))
    op = str(rng.choice(["count", "count", "count-forward", "count-reverse"]))
    # 64 files: at least six prefix bits; the reference's planner stops at 2k - 2 (merylOp-count.C:197), its simple mode
    # uses 6 for k <= 13 (k = 3: no suffix bits at all)
    wp = int(rng.integers(6, max(6, min(2 * k - 2, 17)) + 1))
    env = {}
    if rng.random() < 0.5: env["MB_MAX_PASS_KEYS"] = str(int(rng.choice([4096, 20000, 100000, 1000000])))
    if rng.random() < 0.4: env["MB_GROUP_KEYS"] = str(int(rng.choice([4096, 8192, 65536])))
    if rng.random() < 0.4: env["MB_DENSE_ITEMS"] = str(int(rng.choice([8, 12, 16, 24])))
    if rng.random() < 0.3: env["MB_SORT"] = "2"
    if rng.random() < 0.25: env["MB_MAX_BATCH_BASES"] = str(int(rng.choice([300, 5000, 60000])))
    if rng.random() < 0.15: env["MB_EXTRACT_DENSE"] = "0"
    if rng.random() < 0.2: env["MB_SORT_PINGPONG"] = "1"
    if rng.random() < 0.15: env["MB_COUNT_GROUP"] = "1"
    if rng.random() < 0.15: env["MB_HIST_SINGLE"] = "1"
    if rng.random() < 0.15: env["MB_NO_SUFFIX32"] = "1"
    if rng.random() < 0.1: env["MB_NO_PREFILTER"] = "1"
    if rng.random() < 0.3: env["MB_PARSE_MIN_PIECE"] = str(int(rng.choice([64, 1000])))
    for name in KNOBS:
        os.environ.pop(name, None)
    os.environ.update(env)
    tmp = tempfile.mkdtemp(prefix="mbfuzz_")
    what = {"seed": seed, "k": k, "op": op, "wprefix": wp, "env": env}
    try:
        paths = []
        for i in range(int(rng.integers(1, 4))):
            data, ext = make_input(rng, k)
            p = os.path.join(tmp, "in%d%s" % (i, ext))
            with open(p, "wb") as fh:
                fh.write(data)
            paths.append(p)
        what["bytes"] = sum(os.path.getsize(p) for p in paths)
        ora, out = os.path.join(tmp, "ora.meryl"), os.path.join(tmp, "sim.meryl")
        o = subprocess.run([util.ORACLE_BIN, "k=%d" % k, "wprefix=%d" % wp, op, *paths, "output", ora], capture_output=True, text=True)
        if o.returncode != 0:                    # (an input the reference rejects, a prefix width it cannot have)
            what["skipped"] = o.stderr[-200:]
            return what
        try:
            push = rng.random() < 0.3                       # the merylInput::loadBases seam: the records' bases pushed in pieces
            shards = int(rng.choice([2, 3, 8])) if (not push and rng.random() < 0.25) else 0      # the gpus=N ingest: every shard of every file
            what["push"], what["shards"] = bool(push), shards
            with mb.Counter(k, wp, op, workspace_bytes=int(rng.choice([0, 0, 1 << 28]))) as c:
                for sh in range(shards):
                    for p in paths:
                        c.add_file_shard(p, sh, shards)
                for p in ([] if shards else paths):
                    if not push:
                        c.add_file(p)
                        continue
                    for rec in util.oracle_parse(open(p, "rb").read()).split(b"."):
                        cuts = sorted(set(int(x) for x in rng.integers(0, len(rec) + 1, size=int(rng.integers(0, 4))))) + [len(rec)]
                        at = 0
                        for cut in cuts:
                            c.add_bases(rec[at:cut], end_of_sequence=(cut == len(rec)))
                            at = cut
                c.count()
                st = c.stats()
                c.write_db(out)
            what["kmers"], what["passes"], what["batches"] = int(st.nKmers), int(st.nPasses), int(st.nBatches)
            util.assert_same_db(ora, out)
        except Exception as ex:
            if isinstance(ex, mb.MerylB200Error) and ex.code == -5 and shards and "record boundary" in str(ex):
                # "exact or error": a cut that cannot be proven to be a record start (an empty record's header followed by
                # another header line reads as sequence in the reference's grammar) is refused, not guessed
                what["skipped"] = "shard cut refused"
                return what
            if isinstance(ex, mb.MerylB200Error) and ex.code == -3 and "MB_MAX_PASS_KEYS" in env:
                what["skipped"] = "one prefix bin holds more k-mers than the pass limit drawn for this case"
                return what
            what["failed"] = repr(ex)[:500]
            if keep:
                os.makedirs(keep, exist_ok=True)
                dst = os.path.join(keep, "case_%d" % seed)
                shutil.rmtree(dst, ignore_errors=True)
                shutil.copytree(tmp, dst)
                with open(os.path.join(dst, "case.json"), "w") as fh:
                    json.dump(what, fh)
        return what
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


SET_OPS = ["union", "union-min", "union-max", "union-sum", "intersect", "intersect-min", "intersect-max", "intersect-sum",
           "subtract", "difference", "symmetric-difference"]
VALUE_OPS = [("less-than", 3), ("greater-than", 1), ("at-least", 2), ("at-most", 1), ("equal-to", 2), ("not-equal-to", 1),
             ("increase", 5), ("decrease", 1), ("multiply", 3), ("divide", 2), ("divide-round", 2), ("modulo", 2)]
UNION_KNOBS = ("MB_UNION_MAX_FILES", "MB_RING_BYTES", "MB_STREAM_BLOCK_MAX", "ORA_STREAM_BATCH_MAX", "MB_DECODE")


def one_union_case(seed, keep):
    """A short chain of set / value operations on count databases (and on its own outputs) against the oracle's."""
    import numpy as np
    import util
    import meryl_b200 as mb
    mb.LIB_PATH, mb._lib = SIM_LIB, None
    rng = np.random.default_rng(seed)
    k = int(rng.choice([6, 8, 12, 16, 19, 21, 22, 31, 32, 33, 40, 51, 64]))
    env = {}
    if rng.random() < 0.4: env["MB_UNION_MAX_FILES"] = str(int(rng.choice([1, 3, 7, 20])))
    if rng.random() < 0.3: env["MB_RING_BYTES"] = str(int(rng.choice([4096, 65536])))
    if rng.random() < 0.4:
        env["MB_STREAM_BLOCK_MAX"] = env["ORA_STREAM_BATCH_MAX"] = str(int(rng.choice([1, 2, 7, 64, 500])))
    if rng.random() < 0.5: env["MB_DECODE"] = str(rng.choice(["warp", "auto", "thread"]))
    for name in KNOBS + UNION_KNOBS:
        os.environ.pop(name, None)
    os.environ.update(env)
    tmp = tempfile.mkdtemp(prefix="mbfuzz_")
    what = {"seed": seed, "k": k, "env": env, "steps": []}
    try:
        dbs = []
        shared, _ = make_input(rng, k)
        for i in range(int(rng.integers(1, 4))):
            data, ext = make_input(rng, k)
            p = os.path.join(tmp, "in%d%s" % (i, ext))
            with open(p, "wb") as fh:
                fh.write(data + (b"\n" + shared if rng.random() < 0.6 and ext == ".fa" and shared[:1] == b">" else b""))
            db = os.path.join(tmp, "c%d.meryl" % i)
            o = subprocess.run([util.ORACLE_BIN, "k=%d" % k, "memory=1", str(rng.choice(["count", "count-forward"])), p, "output", db],
                               capture_output=True, text=True)
            if o.returncode != 0:
                what["skipped"] = o.stderr[-200:]
                return what
            dbs.append(db)
        try:
            for step in range(int(rng.integers(1, 4))):
                ora, out = os.path.join(tmp, "o%d.meryl" % step), os.path.join(tmp, "d%d.meryl" % step)
                if rng.random() < 0.7:
                    op = str(rng.choice(SET_OPS))
                    ins = [dbs[int(j)] for j in rng.integers(0, len(dbs), size=int(rng.integers(1, 4)))]
                    what["steps"].append([op, [os.path.basename(x) for x in ins]])
                    util.run_oracle([op, *ins, "output", ora])
                    st = mb.set_operation(op, ins, out)
                else:
                    op, param = VALUE_OPS[int(rng.integers(0, len(VALUE_OPS)))]
                    inp = dbs[int(rng.integers(0, len(dbs)))]
                    what["steps"].append([op, param, os.path.basename(inp)])
                    util.run_oracle([op, str(param), inp, "output", ora])
                    st = mb.value_operation(op, param, inp, out)
                util.assert_same_db(ora, out)
                what["kmers"] = what.get("kmers", 0) + int(st.nInputKmers)
                dbs.append(out)                             # a later step may read this step's output (stream-written blocks)
        except Exception as ex:
            what["failed"] = repr(ex)[:500]
            if keep:
                os.makedirs(keep, exist_ok=True)
                dst = os.path.join(keep, "union_case_%d" % seed)
                shutil.rmtree(dst, ignore_errors=True)
                shutil.copytree(tmp, dst)
                with open(os.path.join(dst, "case.json"), "w") as fh:
                    json.dump(what, fh)
        return what
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


MULTI_KNOBS = ("MB_MULTI_SUFFIX32", "MB_P2P_SCATTER")


def one_lookup_case(seed, keep):
    """`meryl-lookup-b200` (emulated library preloaded) next to the reference's meryl-lookup: -existence and the per-position
    reports on query files with the parser's oddities, one or two databases, value windows; the output files byte for byte."""
    import numpy as np
    import util
    import meryl_b200 as mb
    if not util.have_ref_lookup():
        return {"seed": seed, "skipped": "no reference meryl-lookup"}
    rng = np.random.default_rng(seed)
    for name in KNOBS + UNION_KNOBS + MULTI_KNOBS + ("MB_LOOKUP_BATCH", "MB_LOOKUP_NO_TABLE"):
        os.environ.pop(name, None)
    k = int(rng.choice([8, 13, 16, 21, 22, 31, 32, 33, 51, 64]))
    rep = [["-existence"], ["-bed"], ["-bed-runs"], ["-wig-count"], ["-wig-depth"], ["-include"], ["-exclude"]][int(rng.integers(0, 7))]
    env = dict(os.environ, LD_PRELOAD=SIM_LIB)
    if rng.random() < 0.4: env["MB_LOOKUP_BATCH"] = str(int(rng.choice([300, 5000])))
    if rng.random() < 0.2: env["MB_LOOKUP_NO_TABLE"] = "1"
    if rng.random() < 0.3: env["MB_UNION_MAX_FILES"] = str(int(rng.choice([3, 20])))
    tmp = tempfile.mkdtemp(prefix="mbfuzz_")
    what = {"seed": seed, "k": k, "words": rep, "env": {n: env[n] for n in ("MB_LOOKUP_BATCH", "MB_LOOKUP_NO_TABLE", "MB_UNION_MAX_FILES") if n in env}}
    try:
        dbs, texts = [], []
        for i in range(1 if rep[0] in ("-include", "-exclude") else int(rng.integers(1, 3))):     # (the filters take one database)
            data, ext = make_input(rng, k)
            texts.append((data, ext))
            p = os.path.join(tmp, "in%d%s" % (i, ext))
            with open(p, "wb") as fh:
                fh.write(data)
            db = os.path.join(tmp, "c%d.meryl" % i)
            o = subprocess.run([util.ORACLE_BIN, "k=%d" % k, "memory=1", str(rng.choice(["count", "count-forward"])), p, "output", db], capture_output=True, text=True)
            if o.returncode != 0:
                what["skipped"] = o.stderr[-200:]
                return what
            dbs.append(db)
        q, ext = (make_input(rng, k) if rng.random() < 0.5 else texts[0])      # other sequences, or the database's own (every k-mer found)
        qf = os.path.join(tmp, "query" + ext)
        with open(qf, "wb") as fh:
            fh.write(q)
        words = rep + ["-sequence", qf, "-mers"] + dbs
        if rng.random() < 0.3: words += ["-min", str(int(rng.integers(1, 4)))]
        if rng.random() < 0.3: words += ["-max", str(int(rng.integers(3, 60)))]
        what["words"] = [os.path.basename(w) for w in words]
        r = subprocess.run([util.REF_LOOKUP_BIN] + words + ["-threads", "2", "-output", os.path.join(tmp, "ref.txt")], capture_output=True, text=True)
        if r.returncode != 0:
            what["skipped"] = "the reference fails: " + r.stderr[-150:]
            return what
        o = subprocess.run([mb.LOOKUP_CLI_PATH] + words + ["-output", os.path.join(tmp, "dev.txt")], capture_output=True, text=True, env=env)
        try:
            assert o.returncode == 0, o.stderr[-400:]
            assert open(os.path.join(tmp, "ref.txt"), "rb").read() == open(os.path.join(tmp, "dev.txt"), "rb").read(), "outputs differ"
            what["kmers"] = 1
        except Exception as ex:
            what["failed"] = repr(ex)[:600]
            if keep:
                os.makedirs(keep, exist_ok=True)
                dst = os.path.join(keep, "lookup_case_%d" % seed)
                shutil.rmtree(dst, ignore_errors=True)
                shutil.copytree(tmp, dst)
                with open(os.path.join(dst, "case.json"), "w") as fh:
                    json.dump(what, fh)
        return what
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def one_cli_case(seed, keep):
    """`meryl-b200 <words>` (the emulated library preloaded) next to the reference binary on the same words: option words in
    any order, memory= / n= through the planner (the prefix width decides the bytes), compress, several inputs, gz."""
    import gzip
    import numpy as np
    import util
    import meryl_b200 as mb
    if not util.have_ref():
        return {"seed": seed, "skipped": "no reference binary"}
    rng = np.random.default_rng(seed)
    for name in KNOBS + UNION_KNOBS + MULTI_KNOBS:
        os.environ.pop(name, None)
    k = int(rng.choice([4, 6, 8, 13, 15, 16, 17, 20, 21, 22, 25, 31, 32, 33, 51, 64]))
    op = str(rng.choice(["count", "count", "count-forward", "count-reverse"]))
    opts = ["k=%d" % k, "threads=%d" % int(rng.integers(1, 5))]
    if rng.random() < 0.8: opts.append("memory=%s" % str(rng.choice(["0.5", "1", "2", "4", "8", "0.05"])))
    if rng.random() < 0.4: opts.append("n=%d" % int(rng.choice([1000, 100000, 10000000, 5000000000])))
    if 8 <= k <= 17 and rng.random() < 0.25:                  # (the reference's simple counter allocates 2 * 4^(k-s) bytes: k - s stays small)
        s_len = int(rng.integers(max(1, k - 12), min(k - 6, 7) + 1)) if k - 6 >= max(1, k - 12) else 0
        if s_len:
            opts.append("count-suffix=" + "".join(str(rng.choice(list("ACGTacgt"))) for _ in range(s_len)))
    opts = [opts[int(i)] for i in rng.permutation(len(opts))]
    tmp = tempfile.mkdtemp(prefix="mbfuzz_")
    what = {"seed": seed, "k": k, "words": opts + [op]}
    try:
        ins = []
        for i in range(int(rng.integers(1, 4))):
            data, ext = make_input(rng, k)
            p = os.path.join(tmp, "in%d%s" % (i, ext))
            if rng.random() < 0.25:
                p += ".gz"
                data = gzip.compress(data)
            with open(p, "wb") as fh:
                fh.write(data)
            if rng.random() < 0.2:
                ins.append("compress")
                what["words"].append("compress")
            ins.append(p)
        ref, out = os.path.join(tmp, "ref.meryl"), os.path.join(tmp, "cli.meryl")

        def line(out_dir):
            if not scramble:
                return opts + [op] + ins + ["output", out_dir]
            # the words are order-free (merylCommandBuilder.C:188-549): options between the inputs, the output in front ...
            words = [[w] for w in [op] + ins] + [["output", out_dir]]
            for w, at in zip(opts, places):
                words.insert(int(at * (len(words) + 1)), [w])
            return [x for w in words for x in w]
        scramble, places = rng.random() < 0.5, rng.random(len(opts))
        what["line"] = [os.path.basename(w) for w in line("OUT")]
        r = subprocess.run([util.REF_BIN] + line(ref), capture_output=True, text=True)
        o = subprocess.run([mb.CLI_PATH] + line(out), capture_output=True, text=True, env=dict(os.environ, LD_PRELOAD=SIM_LIB))
        try:
            if r.returncode != 0:
                assert o.returncode != 0, "the reference fails (%s) and meryl-b200 does not" % r.stderr[-150:]
                what["skipped"] = "both refuse"
                return what
            assert o.returncode == 0, o.stderr[-400:]
            util.assert_same_db(ref, out)
            what["kmers"] = 1
            # a second command on the database(s): the readers and the operators' words, outputs compared with the reference's
            sim_env = dict(os.environ, LD_PRELOAD=SIM_LIB)
            for rnd in range(2):
                pick = int(rng.integers(0, 7))
                if pick <= 2:
                    words = [["histogram", ref], ["statistics", ref], ["print", ref]][pick]
                    a = subprocess.run([util.REF_BIN] + words, capture_output=True, text=True)
                    b = subprocess.run([mb.CLI_PATH] + words, capture_output=True, text=True, env=sim_env)
                    what["second"] = words[0]
                    assert a.returncode == b.returncode == 0, b.stderr[-300:]
                    assert (sorted(a.stdout.splitlines()) == sorted(b.stdout.splitlines())) if pick == 2 else (a.stdout == b.stdout), words[0] + " differs"
                    continue
                if k < 6:
                    continue                                  # (the reference's stream writer asserts below k = 6)
                r2, o2 = os.path.join(tmp, "ref2_%d.meryl" % rnd), os.path.join(tmp, "cli2_%d.meryl" % rnd)
                if pick == 3:
                    words = [str(rng.choice(SET_OPS)), ref, ref]
                elif pick == 4:
                    opn, par = VALUE_OPS[int(rng.integers(0, len(VALUE_OPS)))]
                    words = [opn, str(par), ref]
                elif pick == 5:
                    words = [str(rng.choice(["less-than", "greater-than", "at-least", "at-most"])),
                             str(rng.choice(["distinct=0.5", "distinct=0.9", "word-frequency=0.0001", "d=0.2", "f=0.01"])), ref]
                else:
                    words = ["union-sum", "[", "greater-than", "1", ref, "]", "[", "at-most", "3", ref, "]"]
                what["second"] = words[:2]
                a = subprocess.run([util.REF_BIN] + words + ["output", r2], capture_output=True, text=True)
                b = subprocess.run([mb.CLI_PATH] + words + ["output", o2], capture_output=True, text=True, env=sim_env)
                if a.returncode != 0:
                    assert b.returncode != 0, "the reference fails (%s) and meryl-b200 does not" % a.stderr[-150:]
                    continue
                assert b.returncode == 0, b.stderr[-400:]
                util.assert_same_db(r2, o2)
        except Exception as ex:
            what["failed"] = repr(ex)[:600]
            if keep:
                os.makedirs(keep, exist_ok=True)
                dst = os.path.join(keep, "cli_case_%d" % seed)
                shutil.rmtree(dst, ignore_errors=True)
                shutil.copytree(tmp, dst)
                with open(os.path.join(dst, "case.json"), "w") as fh:
                    json.dump(what, fh)
        return what
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def one_multi_case(seed, keep):
    """The collective count of several ranks (fibers of one emulator process, rendezvous copies for NCCL, peer stores for the
    fused exchange) with uneven shares, batches and both routes against the single-process oracle."""
    import ctypes
    import numpy as np
    import util
    import meryl_b200 as mb
    from meryl_b200 import synth
    mb.LIB_PATH, mb._lib = SIM_LIB, None
    rng = np.random.default_rng(seed)
    world = int(rng.choice([2, 2, 3, 4, 5, 7, 8]))
    k = int(rng.choice([8, 15, 16, 19, 21, 22, 31, 32, 33, 51, 64]))
    mode = int(rng.integers(0, 3))
    wp = int(rng.integers(6, max(6, min(2 * k - 2, 15)) + 1))
    env = {}
    if rng.random() < 0.4: env["MB_MAX_BATCH_BASES"] = str(int(rng.choice([20000, 60000, 150000])))
    if rng.random() < 0.4: env["MB_MAX_PASS_KEYS"] = str(int(rng.choice([20000, 100000])))
    if rng.random() < 0.3: env["MB_MULTI_SUFFIX32"] = "0"
    if rng.random() < 0.3: env["MB_P2P_SCATTER"] = "0"
    if rng.random() < 0.3: env["MB_GROUP_KEYS"] = str(int(rng.choice([4096, 65536])))
    if rng.random() < 0.2: env["MB_SORT"] = "2"
    if rng.random() < 0.2: env["MB_EXTRACT_DENSE"] = "0"
    for name in KNOBS + UNION_KNOBS + MULTI_KNOBS:
        os.environ.pop(name, None)
    os.environ.update(env)
    tmp = tempfile.mkdtemp(prefix="mbfuzz_")
    weights = [int(x) for x in rng.integers(0, 5, size=world)]
    if sum(weights) == 0:
        weights[int(rng.integers(0, world))] = 1
    what = {"seed": seed, "world": world, "k": k, "mode": mode, "wprefix": wp, "weights": weights, "env": env}
    try:
        rl = int(rng.choice([max(k + 1, 40), 150, 600]))
        n_reads = int(rng.choice([300, 2000, 6000])) * 150 // rl + 1
        data = synth.reads_fastq_bytes(int(rng.choice([3000, 30000])) + rl, n_reads, rl, float(rng.choice([0.0, 0.004])), seed)
        fq = os.path.join(tmp, "reads.fq")
        with open(fq, "wb") as fh:
            fh.write(data)
        ora, out = os.path.join(tmp, "ora.meryl"), os.path.join(tmp, "sim.meryl")
        util.run_oracle(["k=%d" % k, "wprefix=%d" % wp, ["count", "count-forward", "count-reverse"][mode], fq, "output", ora])
        try:
            recs = data.split(b"@r")[1:]
            cuts = [len(recs) * sum(weights[:r]) // sum(weights) for r in range(world + 1)]
            parts = [b"".join(b"@r" + x for x in recs[cuts[r]:cuts[r + 1]]) for r in range(world)]
            L = mb.lib()
            L.mb_sim_multi_count.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p),
                                             ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)]
            arr = (ctypes.c_char_p * world)(*parts)
            lens = (ctypes.c_uint64 * world)(*[len(p) for p in parts])
            nk, npass = (ctypes.c_uint64 * world)(), (ctypes.c_uint32 * world)()
            os.makedirs(out, exist_ok=True)
            rc = L.mb_sim_multi_count(world, k, mode, wp, arr, lens, out.encode(), nk, npass)
            if rc != 0:
                msg = L.mb_last_error().decode()
                if rc == -3 and "MB_MAX_PASS_KEYS" in env:
                    what["skipped"] = msg[:200]
                    return what
                raise RuntimeError("rc %d: %s" % (rc, msg))
            what["kmers"], what["passes"] = int(sum(nk)), int(max(npass))
            what["batches"] = 2 if "MB_MAX_BATCH_BASES" in env else 1
            util.assert_same_db(ora, out)
        except Exception as ex:
            what["failed"] = repr(ex)[:500]
            if keep:
                os.makedirs(keep, exist_ok=True)
                dst = os.path.join(keep, "multi_case_%d" % seed)
                shutil.rmtree(dst, ignore_errors=True)
                shutil.copytree(tmp, dst)
                with open(os.path.join(dst, "case.json"), "w") as fh:
                    json.dump(what, fh)
        return what
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def worker(args):
    first, step, seconds, keep, what = args
    t0, seed, res = time.time(), first, []
    while time.time() - t0 < seconds:
        res.append({"union": one_union_case, "multi": one_multi_case, "count": one_case, "cli": one_cli_case, "lookup": one_lookup_case}[what](seed, keep))
        seed += step
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--what", default="count", choices=["count", "union", "multi", "cli", "lookup"])
    ap.add_argument("--keep", default=os.path.join(tempfile.gettempdir(), "mbfuzz_failures"))
    a = ap.parse_args()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    subprocess.run(["make", "-s", "-j8", "-C", SIM_DIR], check=True)
    with mp.get_context("spawn").Pool(a.jobs) as pool:
        parts = pool.map(worker, [(a.seed + j, a.jobs, a.seconds, a.keep, a.what) for j in range(a.jobs)])
    res = [r for p in parts for r in p]
    bad = [r for r in res if "failed" in r]
    ran = [r for r in res if "failed" not in r and "skipped" not in r]
    print(json.dumps({"cases": len(res), "compared": len(ran), "skipped": len(res) - len(ran) - len(bad), "failed": len(bad),
                      "kmers": sum(r.get("kmers", 0) for r in ran), "multi_pass": sum(1 for r in ran if r.get("passes", 0) > 1),
                      "multi_batch": sum(1 for r in ran if r.get("batches", 0) > 1)}))
    for r in bad:
        print(json.dumps(r))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

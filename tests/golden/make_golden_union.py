#!/usr/bin/env python
"""Golden fixtures for `meryl union-sum`, written by the UNMODIFIED reference binary
(oracle/_ref/meryl): each case is the union-sum of count dbs that are already golden fixtures here
(their inputs and dbs come from make_golden.py), so only the output db is stored:
    <case>.db.tar.gz   +   manifest_union.json (input cases, sha256 per db file)
Run from the repo root:  python tests/golden/make_golden_union.py
"""
import gzip
import io
import json
import os
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import util  # noqa: E402

CASES = {
    "union_k21_three": ["fasta_k21", "fastq_k21_fwd", "edge_k21"],   # three inputs, different prefix widths, overlapping k-mers
    "union_k51_self":  ["fasta_k51", "fasta_k51"],                # two-word k-mers, every count doubles
}


def main():
    assert util.have_ref(), "build oracle/_ref/meryl first (oracle/build_ref.sh)"
    manifest = {}
    for name, inputs in CASES.items():
        with tempfile.TemporaryDirectory() as t:
            dbs = []
            for i, case in enumerate(inputs):
                sub = os.path.join(t, "in%d" % i)
                os.makedirs(sub)
                dbs.append(util.load_golden(case, sub)[1])
            out = os.path.join(t, "union.meryl")
            util.run_ref(["threads=1", "union-sum", *dbs, "output", out])
            buf = io.BytesIO()
            with tarfile.open(fileobj=buf, mode="w") as tf:
                for f in util.DB_FILES:
                    ti = tf.gettarinfo(os.path.join(out, f), arcname=f)
                    ti.mtime = 0
                    ti.uid = ti.gid = 0
                    ti.uname = ti.gname = ""
                    with open(os.path.join(out, f), "rb") as fh:
                        tf.addfile(ti, fh)
            with open(os.path.join(HERE, name + ".db.tar.gz"), "wb") as raw:
                with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as gz:
                    gz.write(buf.getvalue())
            manifest[name] = {"inputs": inputs, "digest": util.db_digest(out), "files": util.db_manifest(out)}
            print(name, inputs, "->", sum(os.path.getsize(os.path.join(out, f)) for f in util.DB_FILES), "bytes db")
    with open(os.path.join(HERE, "manifest_union.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

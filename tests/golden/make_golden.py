#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the UNMODIFIED
reference binary (oracle/_ref/meryl, built from /root/reference by
oracle/build_ref.sh) on seeded synthetic inputs.

For each case it writes
    <case>.input.gz      the exact input bytes (gzip)
    <case>.db.tar.gz     the 129-file db directory the reference wrote
and one manifest.json with the command line words and a sha256 per db file.
Run from the repo root:  python tests/golden/make_golden.py
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
from meryl_b200 import synth  # noqa: E402

EDGE_FA = (b"\n>rec1 lower and N\nacgtacgtNNNNacgtacgtacgtacgtacgtacgtac\nACGTRYKMacgtacgtacgtacgtacgtacgtacgtacgt\n"
           b">short\nACGT\n>empty\n>after_empty\nACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n"
           b">spaces\nACGT ACGT\tACGTACGT ACGTACGTACGTACGTACGTACGTAC\n"
           b">crlf\r\nACGTTGCAACGTTGCAACGTTGCAACGTTGCA\r\nACGTTGCAACGTTGCAACGTTGCAACGTTGCA\r\n"
           b">noeol\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTGGGGGGGGGGGGGGGGGGGGGGGGG")

CASES = {
    # name: (input file name, bytes, words before the input, )
    "fasta_k21":        ("in.fa", synth.fasta_bytes(3, 40_000, 42),                         ["k=21", "memory=2", "count"]),
    "fastq_k31":        ("in.fq", synth.reads_fastq_bytes(20_000, 2_000, 150, 0.002, 9),    ["k=31", "memory=2", "count"]),
    "fasta_k51":        ("in.fa", synth.fasta_bytes(2, 30_000, 51),                         ["k=51", "memory=2", "count"]),
    "fastq_k21_fwd":    ("in.fq", synth.reads_fastq_bytes(10_000, 1_500, 100, 0.01, 4),     ["k=21", "memory=2", "count-forward"]),
    "edge_k8_simple":   ("in.fa", EDGE_FA,                                                  ["k=8", "memory=2", "count"]),
    "edge_k21":         ("in.fa", EDGE_FA,                                                  ["k=21", "memory=2", "count"]),
    "fasta_k33_rev":    ("in.fa", synth.fasta_bytes(2, 20_000, 33),                         ["k=33", "memory=2", "count-reverse"]),
}


def main():
    assert util.have_ref(), "build oracle/_ref/meryl first (oracle/build_ref.sh)"
    manifest = {}
    for name, (fn, data, words) in CASES.items():
        with tempfile.TemporaryDirectory() as t:
            inp = os.path.join(t, fn)
            with open(inp, "wb") as fh:
                fh.write(data)
            db = os.path.join(t, "db.meryl")
            util.run_ref(words + ["threads=1", fn, "output", "db.meryl"], cwd=t)
            with open(os.path.join(HERE, name + ".input.gz"), "wb") as raw:
                with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as gz:
                    gz.write(data)
            buf = io.BytesIO()
            with tarfile.open(fileobj=buf, mode="w") as tf:
                for f in util.DB_FILES:
                    ti = tf.gettarinfo(os.path.join(db, f), arcname=f)
                    ti.mtime = 0
                    ti.uid = ti.gid = 0
                    ti.uname = ti.gname = ""
                    with open(os.path.join(db, f), "rb") as fh:
                        tf.addfile(ti, fh)
            with open(os.path.join(HERE, name + ".db.tar.gz"), "wb") as raw:
                with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as gz:
                    gz.write(buf.getvalue())
            manifest[name] = {"input": fn, "words": words, "digest": util.db_digest(db), "files": util.db_manifest(db)}
            print(name, len(data), "bytes in ->", sum(os.path.getsize(os.path.join(db, f)) for f in util.DB_FILES), "bytes db")
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

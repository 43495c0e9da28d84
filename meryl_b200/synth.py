"""Seeded synthetic sequence inputs for the parity tests and the bench.

The shapes follow SURVEY.md section 8(d): C1 = 10 x 1 Mbp uniform-random FASTA
wrapped at 80 columns; C2/C3 = 150 bp reads drawn from a uniform-random genome,
both strands, substitution errors, 4-line FASTQ with a constant quality line.
Host-side (numpy) generation is for tests and small benches; the bench's large
workloads are generated on the device by the same recipe (csrc/synth.cu).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b


def random_bases(n: int, rng: np.random.Generator) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def fasta_bytes(nseq: int, seqlen: int, seed: int, width: int = 80, eol: bytes = b"\n") -> bytes:
    rng = np.random.default_rng(seed)
    out = []
    for i in range(nseq):
        s = random_bases(seqlen, rng).tobytes()
        out.append(b">seq%d synthetic" % i + eol)
        for p in range(0, seqlen, width):
            out.append(s[p:p + width] + eol)
    return b"".join(out)


def reads_fastq_bytes(genome_len: int, nreads: int, readlen: int, err: float, seed: int,
                      genome_seed: int | None = None) -> bytes:
    """nreads reads of readlen from a random genome; strand 50/50; substitution rate err."""
    grng = np.random.default_rng(seed if genome_seed is None else genome_seed)
    genome = random_bases(genome_len, grng)
    rng = np.random.default_rng(seed + 1)
    starts = rng.integers(0, genome_len - readlen + 1, size=nreads)
    idx = starts[:, None] + np.arange(readlen)[None, :]
    reads = genome[idx]
    rev = rng.random(nreads) < 0.5
    reads[rev] = _COMP[reads[rev][:, ::-1]]
    if err > 0:
        mask = rng.random(reads.shape) < err
        # substitute with a different base
        cur = np.searchsorted(np.frombuffer(b"ACGT", np.uint8), reads)  # A,C,G,T sorted ascii
        sub = (cur + rng.integers(1, 4, size=reads.shape)) % 4
        reads = np.where(mask, _ACGT[sub], reads)
    qual = b"I" * readlen
    out = bytearray()
    for i in range(nreads):
        out += b"@r%d\n" % i
        out += reads[i].tobytes()
        out += b"\n+\n"
        out += qual
        out += b"\n"
    return bytes(out)

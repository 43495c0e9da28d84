"""The device parser (csrc/parse.cu, kernels K11a-c; experimental, MB_DEVICE_PARSE=1) against the host parser, which is pinned to
the oracle's restatement of the reference's reader (tests/test_host_logic.py).  Shared by the emulator and the GPU tests."""
import ctypes

import numpy as np

import meryl_b200 as mb
from meryl_b200 import synth

EDGE = [
    b"", b"\n\n", b">only header", b">h\nACGT", b">h\nACGT\n", b">h\r\nACGT\r\nTTTT\r\n>i\r\nGG\r\n", b">h\rACGT\rTT\r>i\rGG",
    b">a\n>b\nACGT\n", b">a\n\n\n>b\n\nAC\n\nGT\n\n", b">a\nAC GT\tTT\n", b"@r1\nACGT\n+\nIIII\n@r2\nTTTT\n+r2\n@@@@\n",
    b"@r1\nACGT\n+\n>>>>\n@r2\nTTTT\n+\n++++\n", b"@r1\nACGT\n+\nIIII", b"@r1\nACGT\n+\n", b"@r1\nACGT\n+",
    b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGGGG\r\n+\r\nIIII\r\n", b">x\nACGT+TT\nAA\n+AC\nGGGG\n>y\nCC\n", b">x\nACGT\n+\nAAAA\nCCCC\n",
    b">x\nAC>GT\n@T\n", b"  \n>x desc\nNNNNACGTNNNN\n", b">x\nAC.GT..\n>y\n.\n",
]


def _run(fn, data):
    n = ctypes.c_uint64(0)
    buf = ctypes.create_string_buffer(len(data) + 64)
    rc = fn(data, len(data), buf, len(data) + 64, ctypes.byref(n))
    return rc, buf.raw[: n.value]


def both(data):
    L = mb.lib()
    L.mb_dbg_parse_device.argtypes = L.mb_dbg_parse.argtypes
    return _run(L.mb_dbg_parse, data), _run(L.mb_dbg_parse_device, data)


def case_edges_and_fuzz(monkeypatch, n_fuzz=150):
    """Byte for byte the host parser's stream (runs of '.' included), with the file cut into 32 KB blocks and whole."""
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"ACGTacgtN>@+\n\n\r \tIx.", dtype=np.uint8)
    fuzz = []
    for it in range(n_fuzz):
        n = int(rng.integers(1, 1500 if it % 10 else 70_000))
        fuzz.append((b">" if it % 2 else b">s\n") + alphabet[rng.integers(0, len(alphabet), n)].tobytes())
    big = [synth.reads_fastq_bytes(5_000, 700, 150, 0.01, 2), synth.fasta_bytes(3, 40_000, 3, width=60, eol=b"\r\n"),
           b">long\n" + b"ACGTTGCA" * 20_000 + b"\n"]
    for block in ("32768", None):
        if block:
            monkeypatch.setenv("MB_DPARSE_BLOCK", block)
        else:
            monkeypatch.delenv("MB_DPARSE_BLOCK")
        for data in EDGE + fuzz + big:
            host, dev = both(data)
            assert host == dev, (block, data[:80])
    # a byte that is neither visible nor white space inside a sequence: both refuse
    host, dev = both(b">x\nAC\x01GT\n")
    assert host[0] != 0 and dev[0] == host[0]

"""Shared helpers for the parity tests: run the reference binary, run the C
oracle, compare db directories byte for byte."""
from __future__ import annotations

import ctypes
import filecmp
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "meryl")
REF_LOOKUP_BIN = os.path.join(ORACLE_DIR, "_ref", "meryl-lookup")
ORACLE_BIN = os.path.join(ORACLE_DIR, "meryl_oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

DB_FILES = (["merylIndex"]
            + ["0x%s.merylData" % format(i, "06b") for i in range(64)]
            + ["0x%s.merylIndex" % format(i, "06b") for i in range(64)])


def have_ref() -> bool:
    return os.access(REF_BIN, os.X_OK)


def have_ref_lookup() -> bool:
    return os.access(REF_LOOKUP_BIN, os.X_OK)


def run_ref(args: list[str], cwd: str | None = None, timeout: int = 600) -> str:
    """Run the unmodified reference meryl; returns its stderr (the log)."""
    p = subprocess.run([REF_BIN] + args, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


def run_ref_stdout(args: list[str], timeout: int = 600) -> str:
    """Run the unmodified reference meryl; returns its stdout (print, histogram, statistics)."""
    p = subprocess.run([REF_BIN] + args, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def run_oracle(args: list[str], cwd: str | None = None, timeout: int = 600) -> str:
    p = subprocess.run([ORACLE_BIN] + args, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


def assert_same_db(a: str, b: str) -> None:
    """All 129 files present in both and byte-identical."""
    la, lb = sorted(os.listdir(a)), sorted(os.listdir(b))
    assert la == sorted(DB_FILES), "unexpected file set in %s: %s" % (a, set(la) ^ set(DB_FILES))
    assert lb == sorted(DB_FILES), "unexpected file set in %s: %s" % (b, set(lb) ^ set(DB_FILES))
    match, mismatch, errors = filecmp.cmpfiles(a, b, DB_FILES, shallow=False)
    assert not mismatch and not errors, "db differs: %s %s" % (mismatch[:5], errors[:5])


def db_manifest(d: str) -> dict[str, str]:
    out = {}
    for f in DB_FILES:
        with open(os.path.join(d, f), "rb") as fh:
            out[f] = hashlib.sha256(fh.read()).hexdigest()
    return out


def db_digest(d: str) -> str:
    """One sha256 over name+content of the 129 files, in DB_FILES order."""
    h = hashlib.sha256()
    for f in DB_FILES:
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode())
            h.update(hashlib.sha256(fh.read()).digest())
    return h.hexdigest()


def ref_wprefix(k: int, n: int, memory_gb: float) -> int:
    """Ask the reference which prefix width it would pick (config-only run, '-C')."""
    # A count needs an input word; /dev/null is rejected, so use n= with a tiny real file.
    import tempfile
    with tempfile.TemporaryDirectory() as t:
        fa = os.path.join(t, "x.fa")
        with open(fa, "w") as fh:
            fh.write(">x\nACGT\n")
        log = run_ref(["-C", "k=%d" % k, "n=%d" % n, "memory=%s" % memory_gb, "count", fa, "output", os.path.join(t, "o")])
    if "Configured simple mode" in log:
        # meryl/merylOp-countSimple.C:196-199; the db bytes of the k-sweep tests pin this formula.
        return 2 * k - min(20, 2 * k - 6)
    for line in log.splitlines():
        if "Best Value!" in line:
            return int(line.split()[0])
    raise AssertionError("no Best Value in reference log:\n" + log[-1500:])


_lib = None


def oracle_lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(ORACLE_LIB)
        L.ora_plan.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                               ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint64),
                               ctypes.POINTER(ctypes.c_uint32)]
        L.ora_plan.restype = ctypes.c_int
        L.ora_buf_new.restype = ctypes.c_void_p
        L.ora_buf_free.argtypes = [ctypes.c_void_p]
        L.ora_buf_ptr.argtypes = [ctypes.c_void_p]
        L.ora_buf_ptr.restype = ctypes.c_void_p
        L.ora_buf_len.argtypes = [ctypes.c_void_p]
        L.ora_buf_len.restype = ctypes.c_size_t
        L.ora_parse.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
        L.ora_parse.restype = ctypes.c_int
        L.ora_extract64.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32,
                                    ctypes.c_void_p, ctypes.c_void_p]
        L.ora_extract64.restype = ctypes.c_uint64
        L.ora_encode_block64.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ora_encode_block64.restype = ctypes.c_size_t
        _lib = L
    return _lib


def oracle_plan(k: int, n: int, mem_bytes: int) -> int:
    L = oracle_lib()
    wp = ctypes.c_uint32(0)
    rc = L.ora_plan(k, n, mem_bytes, ctypes.byref(wp), None, None)
    assert rc == 0
    return wp.value


def oracle_parse(data: bytes) -> bytes:
    L = oracle_lib()
    b = L.ora_buf_new()
    try:
        rc = L.ora_parse(data, len(data), b)
        if rc:
            raise ValueError("oracle parse error %d" % rc)
        return ctypes.string_at(L.ora_buf_ptr(b), L.ora_buf_len(b))
    finally:
        L.ora_buf_free(b)


def oracle_extract(bases: bytes, k: int, mode: int = 0):
    """-> (hi, lo) numpy uint64 arrays of k-mers in stream order."""
    import numpy as np
    L = oracle_lib()
    n = L.ora_extract64(bases, len(bases), k, mode, None, None)
    hi = np.zeros(n, dtype=np.uint64)
    lo = np.zeros(n, dtype=np.uint64)
    if n:
        L.ora_extract64(bases, len(bases), k, mode, hi.ctypes.data, lo.ctypes.data)
    return hi, lo


def oracle_encode_block(w_suffix: int, prefix: int, suf_hi, suf_lo, values) -> bytes:
    import numpy as np
    L = oracle_lib()
    n = len(suf_lo)
    suf_lo = np.ascontiguousarray(suf_lo, dtype=np.uint64)
    suf_hi = np.ascontiguousarray(suf_hi if suf_hi is not None else np.zeros(n, np.uint64), dtype=np.uint64)
    values = np.ascontiguousarray(values, dtype=np.uint32)
    b = L.ora_buf_new()
    try:
        L.ora_encode_block64(w_suffix, prefix, n, suf_hi.ctypes.data, suf_lo.ctypes.data, values.ctypes.data, b)
        return ctypes.string_at(L.ora_buf_ptr(b), L.ora_buf_len(b))
    finally:
        L.ora_buf_free(b)


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_cases() -> list[str]:
    import json
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as fh:
        return sorted(json.load(fh).keys())


def load_golden(name: str, tmp: str):
    """Unpacks a golden case into tmp; returns (input path, reference db dir, command words, manifest entry)."""
    import gzip
    import io
    import json
    import tarfile
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as fh:
        ent = json.load(fh)[name]
    inp = os.path.join(tmp, ent["input"])
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".input.gz"), "rb") as g, open(inp, "wb") as o:
        o.write(g.read())
    db = os.path.join(tmp, "golden.meryl")
    os.makedirs(db)
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".db.tar.gz"), "rb") as g:
        with tarfile.open(fileobj=io.BytesIO(g.read())) as tf:
            tf.extractall(db, filter="data")
    assert db_digest(db) == ent["digest"]
    return inp, db, ent["words"], ent


def union_golden_cases() -> list[str]:
    import json
    with open(os.path.join(GOLDEN_DIR, "manifest_union.json")) as fh:
        return sorted(json.load(fh))


def load_union_golden(name: str, tmp: str):
    """Unpacks a union-sum golden case: returns (list of input db dirs, reference output db dir)."""
    import gzip
    import io
    import json
    import tarfile
    with open(os.path.join(GOLDEN_DIR, "manifest_union.json")) as fh:
        ent = json.load(fh)[name]
    dbs = []
    for i, case in enumerate(ent["inputs"]):
        sub = os.path.join(tmp, "in%d" % i)
        os.makedirs(sub)
        dbs.append(load_golden(case, sub)[1])
    out = os.path.join(tmp, "golden_union.meryl")
    os.makedirs(out)
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".db.tar.gz"), "rb") as g:
        with tarfile.open(fileobj=io.BytesIO(g.read())) as tf:
            tf.extractall(out, filter="data")
    assert db_digest(out) == ent["digest"]
    return dbs, out

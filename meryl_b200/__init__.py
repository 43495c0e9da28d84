"""meryl_b200 -- host-side Python mirror of the B200 `meryl count` path.

The product is the C-ABI shared library ``libmeryl_b200.so`` (built in tree by
``meryl_b200/csrc/Makefile``; ABI in ``include/meryl_b200.h``) and the
``meryl-b200`` command line tool.  This module only binds that ABI with ctypes so
that tests, ``bench.py`` and Python callers can drive it; it contains no compute
and no fallback: if the library or a CUDA device is missing, calls raise.

The reference has no Python surface; names follow its command words
(``count``, ``count-forward``, ``count-reverse``, ``k=``, ``n=``, ``memory=``,
``output``; src/meryl/merylCommandBuilder.C:188-326, :346-348, :441-461).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmeryl_b200.so")
CLI_PATH = os.path.join(_HERE, "meryl-b200")
LOOKUP_CLI_PATH = os.path.join(_HERE, "meryl-lookup-b200")

MODE_CANONICAL, MODE_FORWARD, MODE_REVERSE = 0, 1, 2
_MODES = {"count": MODE_CANONICAL, "count-forward": MODE_FORWARD, "count-reverse": MODE_REVERSE}

FLAG_KEEP_COUNTED = 1
FLAG_NO_IMAGES = 2
FLAG_COMPRESS = 4


class MerylB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("meryl_b200 error %d: %s" % (code, msg))
        self.code = code


class _Config(ctypes.Structure):
    _fields_ = [("k", ctypes.c_uint32), ("mode", ctypes.c_uint32), ("wPrefix", ctypes.c_uint32),
                ("device", ctypes.c_int32), ("workspaceBytes", ctypes.c_uint64),
                ("rank", ctypes.c_uint32), ("worldSize", ctypes.c_uint32),
                ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [("nBases", ctypes.c_uint64), ("nKmers", ctypes.c_uint64), ("nDistinct", ctypes.c_uint64),
                ("nUnique", ctypes.c_uint64), ("nTotal", ctypes.c_uint64), ("dbBytes", ctypes.c_uint64),
                ("nPasses", ctypes.c_uint32), ("nLaunches", ctypes.c_uint32),
                ("h2dBytes", ctypes.c_uint64), ("d2hBytes", ctypes.c_uint64),
                ("msPack", ctypes.c_float), ("msHistogram", ctypes.c_float), ("msExtract", ctypes.c_float),
                ("msSortHist", ctypes.c_float), ("msSort", ctypes.c_float), ("msCount", ctypes.c_float),
                ("msEncode", ctypes.c_float), ("msExchange", ctypes.c_float), ("msTotal", ctypes.c_float),
                ("sortPasses", ctypes.c_uint32), ("keyBytes", ctypes.c_uint32), ("sortLaunchKeys", ctypes.c_uint64),
                ("nBatches", ctypes.c_uint32), ("msMerge", ctypes.c_float)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class UnionStats(ctypes.Structure):
    _fields_ = [("nInputKmers", ctypes.c_uint64), ("nOutputKmers", ctypes.c_uint64),
                ("nUnique", ctypes.c_uint64), ("nDistinct", ctypes.c_uint64), ("nTotal", ctypes.c_uint64),
                ("inBytes", ctypes.c_uint64), ("outBytes", ctypes.c_uint64),
                ("h2dBytes", ctypes.c_uint64), ("d2hBytes", ctypes.c_uint64),
                ("nInputs", ctypes.c_uint32), ("nGroups", ctypes.c_uint32), ("nLaunches", ctypes.c_uint32), ("k", ctypes.c_uint32),
                ("msDecode", ctypes.c_float), ("msMerge", ctypes.c_float), ("msEncode", ctypes.c_float), ("msTotal", ctypes.c_float)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib() -> ctypes.CDLL:
    """The C-ABI library; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MerylB200Error(-2, "%s not built; run `make -C meryl_b200/csrc` (no CPU fallback exists)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        u8p, u32p, u64p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint64)
        L.mb_last_error.restype = ctypes.c_char_p
        L.mb_plan.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, u32p, ctypes.POINTER(ctypes.c_int), u64p, u32p]
        L.mb_estimate_kmers_in_file.argtypes = [ctypes.c_char_p]
        L.mb_estimate_kmers_in_file.restype = ctypes.c_uint64
        L.mb_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(ctypes.c_void_p)]
        L.mb_destroy.argtypes = [ctypes.c_void_p]
        L.mb_destroy.restype = None
        L.mb_add_bases.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        L.mb_add_bases_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.mb_add_file.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mb_add_file_bytes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.mb_add_file_shard.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32]
        L.mb_comm_unique_id.argtypes = [ctypes.c_void_p]
        L.mb_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.mb_file_owner_of.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.mb_file_owner_of.restype = ctypes.c_uint32
        L.mb_set_output_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.mb_set_output_dir.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mb_set_input_compression.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.mb_count.argtypes = [ctypes.c_void_p]
        L.mb_get_image.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p), u64p]
        L.mb_write_db.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mb_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
        L.mb_get_counted.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.POINTER(ctypes.c_void_p), u64p]
        L.mb_get_histogram.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), u64p]
        L.mb_merge_histogram.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.mb_count_files.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(ctypes.c_char_p), ctypes.c_uint32,
                                     ctypes.c_char_p, ctypes.POINTER(Stats)]
        L.mb_dbg_extract_sorted.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, u64p]
        L.mb_dbg_sort.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32]
        L.mb_dbg_sort32.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32]
        L.mb_dbg_encode_block.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_uint64, u64p]
        L.mb_dbg_parse.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, u64p]
        L.mb_synth_record_bytes.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.mb_synth_record_bytes.restype = ctypes.c_uint64
        L.mb_synth_reads_device.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                            ctypes.c_uint64, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise MerylB200Error(rc, (lib().mb_last_error() or b"").decode(errors="replace"))


@dataclass
class Plan:
    w_prefix: int
    simple: bool
    mem_used: int
    n_batches: int


def plan(k: int, n_estimate: int, memory_bytes: int) -> Plan:
    """== configureCounting (src/meryl/merylOp-count.C:300-403): the prefix width the reference picks."""
    wp, simple = ctypes.c_uint32(0), ctypes.c_int(0)
    used, nb = ctypes.c_uint64(0), ctypes.c_uint32(0)
    _check(lib().mb_plan(k, n_estimate, memory_bytes, ctypes.byref(wp), ctypes.byref(simple), ctypes.byref(used), ctypes.byref(nb)))
    return Plan(wp.value, bool(simple.value), used.value, nb.value)


def estimate_kmers_in_file(path: str) -> int:
    return lib().mb_estimate_kmers_in_file(os.fsencode(path))


class Counter:
    """One counting job on one GPU (== one merylOperation::countThreads call)."""

    def __init__(self, k: int, w_prefix: int, operation: str = "count", device: int = -1,
                 workspace_bytes: int = 0, flags: int = 0, rank: int = 0, world_size: int = 1):
        if operation not in _MODES:
            raise ValueError("operation must be one of %s" % sorted(_MODES))
        self.k, self.w_prefix = k, w_prefix
        cfg = _Config(k, _MODES[operation], w_prefix, device, workspace_bytes, rank, world_size, flags, 0)
        self._h = ctypes.c_void_p()
        _check(lib().mb_create(ctypes.byref(cfg), ctypes.byref(self._h)))

    def close(self) -> None:
        if self._h:
            lib().mb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- input (merylInput::loadBases seam) ----
    def add_bases(self, bases: bytes, end_of_sequence: bool = True) -> None:
        _check(lib().mb_add_bases(self._h, bases, len(bases), int(end_of_sequence)))

    def add_bases_ptr(self, ptr: int, n: int, end_of_sequence: bool = True) -> None:
        _check(lib().mb_add_bases(self._h, ctypes.c_void_p(ptr), n, int(end_of_sequence)))

    def add_bases_device(self, dev_ptr: int, n: int) -> None:
        _check(lib().mb_add_bases_device(self._h, ctypes.c_void_p(dev_ptr), n))

    def add_file(self, path: str) -> None:
        _check(lib().mb_add_file(self._h, os.fsencode(path)))

    def add_file_shard(self, path: str, shard: int, n_shards: int) -> None:
        _check(lib().mb_add_file_shard(self._h, os.fsencode(path), shard, n_shards))

    def add_file_bytes(self, data: bytes) -> None:
        _check(lib().mb_add_file_bytes(self._h, data, len(data)))

    def add_file_bytes_ptr(self, ptr: int, n: int) -> None:
        _check(lib().mb_add_file_bytes(self._h, ctypes.c_void_p(ptr), n))

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _check(lib().mb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes) -> None:
        _check(lib().mb_comm_init(self._h, unique_id))

    # ---- run / results ----
    def set_output_buffer(self, host_ptr: int, n_bytes: int) -> None:
        """Caller-owned (ideally pinned) host memory for the 64 data-file images."""
        _check(lib().mb_set_output_buffer(self._h, ctypes.c_void_p(host_ptr), n_bytes))

    def set_input_compression(self, on: bool) -> None:
        """Homopolymer compression for the inputs added from now on (the CLI's `compress` word)."""
        _check(lib().mb_set_input_compression(self._h, 1 if on else 0))

    def set_output_dir(self, out_dir: str) -> None:
        """Before count(): stream the data files into out_dir while counting (then write_db(out_dir) adds the indexes)."""
        _check(lib().mb_set_output_dir(self._h, os.fsencode(out_dir)))

    def count(self) -> None:
        _check(lib().mb_count(self._h))

    def write_db(self, out_dir: str) -> None:
        _check(lib().mb_write_db(self._h, os.fsencode(out_dir)))

    def image(self, which: int) -> bytes:
        p, n = ctypes.c_void_p(), ctypes.c_uint64(0)
        _check(lib().mb_get_image(self._h, which, ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p, n.value) if n.value else b""

    def stats(self) -> Stats:
        st = Stats()
        _check(lib().mb_get_stats(self._h, ctypes.byref(st)))
        return st

    def counted(self):
        """(lo, hi, count) numpy arrays of the distinct k-mers (needs FLAG_KEEP_COUNTED)."""
        import numpy as np
        lo, hi, cnt, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64(0)
        _check(lib().mb_get_counted(self._h, ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(cnt), ctypes.byref(n)))
        m = n.value

        def arr(p, t):
            if not m or not p.value:
                return np.zeros(m, dtype=t)
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(np.ctypeslib.as_ctypes_type(t))), shape=(m,)).copy()
        return arr(lo, np.uint64), arr(hi, np.uint64), arr(cnt, np.uint32)

    def histogram(self):
        import numpy as np
        v, o, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64(0)
        _check(lib().mb_get_histogram(self._h, ctypes.byref(v), ctypes.byref(o), ctypes.byref(n)))
        m = n.value
        if not m:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
        f = lambda p: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(m,)).copy()
        return f(v), f(o)

    def merge_histogram(self, values, occurrences) -> None:
        import numpy as np
        values = np.ascontiguousarray(values, dtype=np.uint64)
        occurrences = np.ascontiguousarray(occurrences, dtype=np.uint64)
        _check(lib().mb_merge_histogram(self._h, values.ctypes.data, occurrences.ctypes.data, len(values)))


def file_owner(file: int, world_size: int) -> int:
    """Rank that owns data file `file` (0..63) when the db is sharded over world_size GPUs."""
    return lib().mb_file_owner_of(file, world_size)


def synth_record_bytes(read_len: int, fmt: int) -> int:
    return lib().mb_synth_record_bytes(read_len, fmt)


def synth_reads_device(dev_ptr: int, first_read: int, n_reads: int, read_len: int, genome_len: int,
                       err: float, seed: int, fmt: int, stream: int = 0) -> None:
    """Fill device memory with seeded synthetic reads (fmt 0: "<bases>." records, fmt 1: FASTQ)."""
    _check(lib().mb_synth_reads_device(ctypes.c_void_p(dev_ptr), first_read, n_reads, read_len, genome_len,
                                       err, seed, fmt, ctypes.c_void_p(stream)))


def count(inputs, output: str, k: int, operation: str = "count", memory_gb: float | None = None,
          n: int | None = None, device: int = -1) -> Stats:
    """`meryl k=K [n=N] [memory=GB] count <inputs> output <output>` on one GPU."""
    inputs = [inputs] if isinstance(inputs, (str, os.PathLike)) else list(inputs)
    if n is None or n == 0:
        n = sum(estimate_kmers_in_file(os.fspath(p)) for p in inputs)
    if memory_gb is None:
        memory_bytes = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    else:
        memory_bytes = int(memory_gb * 1024 * 1024 * 1024)
    pl = plan(k, n, memory_bytes)
    with Counter(k, pl.w_prefix, operation, device) as c:
        c.set_output_dir(os.fspath(output))
        for p in inputs:
            c.add_file(os.fspath(p))
        c.count()
        c.write_db(os.fspath(output))
        return c.stats()


def union_sum(inputs, output: str, device: int = -1, workspace_bytes: int = 0) -> UnionStats:
    """`meryl union-sum <inputs> output <output>` on one GPU (include/meryl_b200.h: mb_union_sum)."""
    inputs = [os.fsencode(os.fspath(p)) for p in inputs]
    L = lib()
    L.mb_union_sum.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int32,
                               ctypes.c_uint64, ctypes.POINTER(UnionStats)]
    arr = (ctypes.c_char_p * len(inputs))(*inputs)
    st = UnionStats()
    _check(L.mb_union_sum(arr, len(inputs), os.fsencode(os.fspath(output)), device, workspace_bytes, ctypes.byref(st)))
    return st


SET_OPERATIONS = ("union", "union-min", "union-max", "union-sum", "intersect", "intersect-min", "intersect-max", "intersect-sum",
                  "subtract", "difference", "symmetric-difference")


def set_operation(op: str, inputs, output: str, device: int = -1, workspace_bytes: int = 0) -> UnionStats:
    """`meryl <op> <inputs> output <output>` for the merge family of operators (include/meryl_b200.h: mb_set_operation)."""
    inputs = [os.fsencode(os.fspath(p)) for p in inputs]
    L = lib()
    L.mb_set_operation.argtypes = [ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p), ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int32,
                                   ctypes.c_uint64, ctypes.POINTER(UnionStats)]
    arr = (ctypes.c_char_p * len(inputs))(*inputs)
    st = UnionStats()
    _check(L.mb_set_operation(SET_OPERATIONS.index(op), arr, len(inputs), os.fsencode(os.fspath(output)), device, workspace_bytes, ctypes.byref(st)))
    return st


VALUE_OPERATIONS = ("less-than", "greater-than", "at-least", "at-most", "equal-to", "not-equal-to",
                    "increase", "decrease", "multiply", "divide", "divide-round", "modulo")


def value_operation(op: str, param: int, inp, output: str, device: int = -1, workspace_bytes: int = 0) -> UnionStats:
    """`meryl <op> <param> <input> output <output>` for the one-input filters / arithmetic (mb_value_operation)."""
    L = lib()
    L.mb_value_operation.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32,
                                     ctypes.c_uint64, ctypes.POINTER(UnionStats)]
    st = UnionStats()
    _check(L.mb_value_operation(11 + VALUE_OPERATIONS.index(op), param, os.fsencode(os.fspath(inp)), os.fsencode(os.fspath(output)),
                                device, workspace_bytes, ctypes.byref(st)))
    return st


def db_threshold(db, distinct: float | None = None, word_frequency: float | None = None) -> int:
    """The filters' threshold given as `distinct=` / `word-frequency=` (merylOperation::initializeThreshold): mb_db_threshold."""
    L = lib()
    L.mb_db_threshold.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_uint64)]
    t = ctypes.c_uint64()
    _check(L.mb_db_threshold(os.fsencode(os.fspath(db)), -1.0 if distinct is None else distinct,
                             -1.0 if word_frequency is None else word_frequency, ctypes.byref(t)))
    return t.value


def keep_suffix(inp, suffix: str, output: str, device: int = -1, workspace_bytes: int = 0) -> UnionStats:
    """`count-suffix=<bases>`: the k-mers of a count db that end with `suffix`, in the simple counter's geometry (mb_keep_suffix)."""
    L = lib()
    L.mb_keep_suffix.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_uint64, ctypes.POINTER(UnionStats)]
    st = UnionStats()
    _check(L.mb_keep_suffix(os.fsencode(os.fspath(inp)), suffix.encode(), os.fsencode(os.fspath(output)), device, workspace_bytes, ctypes.byref(st)))
    return st


def lookup(db, kmer_lo, kmer_hi=None, min_value: int = 0, max_value: int = 0xffffffff, device: int = -1, workspace_bytes: int = 0):
    """merylExactLookup::value for an array of k-mers (mb_lookup): numpy uint64 arrays in (2 bits per base, first base in the
    top bits, A=0 C=1 T=2 G=3; kmer_hi for k > 32), numpy uint32 values out (0 = absent)."""
    import numpy as np
    lo = np.ascontiguousarray(kmer_lo, dtype=np.uint64)
    hi = None if kmer_hi is None else np.ascontiguousarray(kmer_hi, dtype=np.uint64)
    out = np.zeros(lo.size, dtype=np.uint32)
    L = lib()
    L.mb_lookup.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                            ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64]
    _check(L.mb_lookup(os.fsencode(os.fspath(db)), min_value, max_value, lo.ctypes.data, None if hi is None else hi.ctypes.data,
                       lo.size, out.ctypes.data, device, workspace_bytes))
    return out


def all_kmers(db, device: int = -1, workspace_bytes: int = 0):
    """Every (k-mer lo, k-mer hi or None, value) of a database in ascending order (mb_for_each_kmer), as numpy arrays."""
    import numpy as np
    los, his, vals = [], [], []
    CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                          ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint64)

    def piece(_arg, _k, lo, hi, val, n):
        los.append(np.ctypeslib.as_array(lo, shape=(n,)).copy())
        if hi:
            his.append(np.ctypeslib.as_array(hi, shape=(n,)).copy())
        vals.append(np.ctypeslib.as_array(val, shape=(n,)).copy())
        return 0

    L = lib()
    L.mb_for_each_kmer.argtypes = [ctypes.c_char_p, CB, ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64]
    _check(L.mb_for_each_kmer(os.fsencode(os.fspath(db)), CB(piece), None, device, workspace_bytes))
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt)
    return cat(los, np.uint64), (cat(his, np.uint64) if his else None), cat(vals, np.uint32)


def existence(db, sequences, min_value: int = 0, max_value: int = 0xffffffff, skip_first: int = 0, device: int = -1, workspace_bytes: int = 0):
    """meryl-lookup -existence for a list of sequences (bytes): (nTotal[], nFound[], nDbKmers) (mb_existence)."""
    import numpy as np
    starts = np.zeros(len(sequences) + 1, dtype=np.uint64)
    if sequences:
        starts[1:] = np.cumsum([len(x) for x in sequences], dtype=np.uint64)
    blob = b"".join(sequences)
    tot = np.zeros(len(sequences), dtype=np.uint64)
    fnd = np.zeros(len(sequences), dtype=np.uint64)
    ndb = ctypes.c_uint64(0)
    L = lib()
    L.mb_existence.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64,
                               ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32, ctypes.c_uint64]
    _check(L.mb_existence(os.fsencode(os.fspath(db)), min_value, max_value, blob, starts.ctypes.data, len(sequences), skip_first,
                          tot.ctypes.data, fnd.ctypes.data, ctypes.byref(ndb), device, workspace_bytes))
    return tot, fnd, ndb.value


def position_lookup(db, sequences, min_value: int = 0, max_value: int = 0xffffffff, values: bool = True, device: int = -1, workspace_bytes: int = 0):
    """Per position of the concatenated sequences: flags (bit 0 a k-mer starts here, bit 1 found) and value(fmer) + value(rmer)
    (mb_position_lookup; meryl-lookup -bed / -wig-count / -wig-depth)."""
    import numpy as np
    starts = np.zeros(len(sequences) + 1, dtype=np.uint64)
    if sequences:
        starts[1:] = np.cumsum([len(x) for x in sequences], dtype=np.uint64)
    blob = b"".join(sequences)
    flags = np.zeros(len(blob), dtype=np.uint8)
    vals = np.zeros(len(blob), dtype=np.uint32) if values else None
    L = lib()
    L.mb_position_lookup.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64]
    _check(L.mb_position_lookup(os.fsencode(os.fspath(db)), min_value, max_value, blob, starts.ctypes.data, len(sequences),
                                flags.ctypes.data, None if vals is None else vals.ctypes.data, device, workspace_bytes))
    return flags, vals


class Table:
    """A database decoded into device memory once and asked many times (mb_table_*; merylExactLookup as an object)."""

    def __init__(self, db, min_value: int = 0, max_value: int = 0xffffffff, device: int = -1, workspace_bytes: int = 0):
        L = lib()
        L.mb_table_open.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int32, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]
        L.mb_table_close.argtypes = [ctypes.c_void_p]
        L.mb_table_close.restype = None
        L.mb_table_k.argtypes = [ctypes.c_void_p]
        L.mb_table_k.restype = ctypes.c_uint32
        L.mb_table_kmers.argtypes = [ctypes.c_void_p]
        L.mb_table_kmers.restype = ctypes.c_uint64
        L.mb_table_query.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.mb_table_values.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        self._h = ctypes.c_void_p()
        _check(L.mb_table_open(os.fsencode(os.fspath(db)), min_value, max_value, device, workspace_bytes, ctypes.byref(self._h)))
        self.k = L.mb_table_k(self._h)
        self.kmers = L.mb_table_kmers(self._h)

    def close(self) -> None:
        if self._h:
            lib().mb_table_close(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, sequences, skip_first: int = 0, positions: bool = False, values: bool = False):
        """(nTotal[], nFound[], flags[] or None, values[] or None) for a list of sequences (bytes)."""
        import numpy as np
        starts = np.zeros(len(sequences) + 1, dtype=np.uint64)
        if sequences:
            starts[1:] = np.cumsum([len(x) for x in sequences], dtype=np.uint64)
        blob = b"".join(sequences)
        tot = np.zeros(len(sequences), dtype=np.uint64)
        fnd = np.zeros(len(sequences), dtype=np.uint64)
        flags = np.zeros(len(blob), dtype=np.uint8) if positions or values else None
        vals = np.zeros(len(blob), dtype=np.uint32) if values else None
        _check(lib().mb_table_query(self._h, blob, starts.ctypes.data, len(sequences), skip_first, tot.ctypes.data, fnd.ctypes.data,
                                    None if flags is None else flags.ctypes.data, None if vals is None else vals.ctypes.data))
        return tot, fnd, flags, vals

    def values(self, kmer_lo, kmer_hi=None):
        import numpy as np
        lo = np.ascontiguousarray(kmer_lo, dtype=np.uint64)
        hi = None if kmer_hi is None else np.ascontiguousarray(kmer_hi, dtype=np.uint64)
        out = np.zeros(lo.size, dtype=np.uint32)
        _check(lib().mb_table_values(self._h, lo.ctypes.data, None if hi is None else hi.ctypes.data, lo.size, out.ctypes.data))
        return out

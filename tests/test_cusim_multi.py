"""The multi-GPU path on the CPU: `world` ranks run as fibers inside the cusim emulator (tests/cusim), the NCCL collectives
of csrc/exchange.cu are replaced by in-process rendezvous copies with the same semantics, and the directory written by all
ranks must be byte-identical to the single-process oracle db -- the CPU twin of tests/test_gpu_multi.py, which needs
2-8 GPUs."""
import ctypes
import os
import subprocess

import pytest

import util
import meryl_b200 as mb
from meryl_b200 import dist as mbd
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


def multi_count(world, k, w_prefix, data: bytes, out_dir: str, mode: int = 0, weights=None):
    recs = data.split(b"@r")[1:]                          # synthetic FASTQ: every record starts with "@r"
    parts = []
    if weights:                                           # uneven shares (rank r gets weights[r] parts of sum(weights))
        cuts = [len(recs) * sum(weights[:r]) // sum(weights) for r in range(world + 1)]
        parts = [b"".join(b"@r" + x for x in recs[cuts[r]:cuts[r + 1]]) for r in range(world)]
    for r in range(world if not weights else 0):
        first, cnt = mbd.shard_range(len(recs), r, world)
        parts.append(b"".join(b"@r" + x for x in recs[first:first + cnt]))
    L = mb.lib()
    L.mb_sim_multi_count.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p),
                                     ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)]
    arr = (ctypes.c_char_p * world)(*parts)
    lens = (ctypes.c_uint64 * world)(*[len(p) for p in parts])
    nk = (ctypes.c_uint64 * world)()
    npass = (ctypes.c_uint32 * world)()
    os.makedirs(out_dir, exist_ok=True)
    rc = L.mb_sim_multi_count(world, k, mode, w_prefix, arr, lens, out_dir.encode(), nk, npass)
    multi_key_bytes.world = world
    assert rc == 0, L.mb_last_error().decode()
    return list(nk), list(npass)


def multi_key_bytes():
    """Key bytes every rank of the last multi_count() sorted with."""
    L = mb.lib()
    L.mb_sim_last_key_bytes.restype = ctypes.c_uint32
    return [L.mb_sim_last_key_bytes(r) for r in range(multi_key_bytes.world)]


@pytest.mark.parametrize("world,k", [(2, 21), (4, 21), (8, 31), (2, 51), (3, 21), (5, 31), (6, 21), (7, 51)])   # (3, 5, 6, 7 do not divide 64: a short last slot)
def test_sim_multi_rank_db_bit_exact(tmp_path, world, k):
    n_reads, rl = 6_000, 150
    data = synth.reads_fastq_bytes(20_000, n_reads, rl, 0.003, 17 + k)
    fq = tmp_path / "reads.fq"
    fq.write_bytes(data)
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=12", "count", str(fq), "output", ora])
    nk, npass = multi_count(world, k, 12, data, out)
    util.assert_same_db(ora, out)
    assert sum(nk) == n_reads * (rl - k + 1)                # every k-mer is counted by exactly one rank
    assert max(nk) < 1.6 * sum(nk) / world                  # and the file ownership map balances the ranks


@pytest.mark.parametrize("world,k,cap", [(2, 21, 60_000), (4, 31, 40_000), (2, 51, 80_000)])
def test_sim_multi_rank_multi_pass(tmp_path, monkeypatch, world, k, cap):
    """100 Gbp on 2 or 4 GPUs does not fit one pass (64-bit keys): passes over groups of file slots, one exchange per pass.
    MB_MAX_PASS_KEYS caps what a rank may send or receive per pass, so a small input takes the same road.  The NCCL route
    with full keys (both defaults of round 2 switched off)."""
    monkeypatch.setenv("MB_MAX_PASS_KEYS", str(cap))
    monkeypatch.setenv("MB_MULTI_SUFFIX32", "0")
    monkeypatch.setenv("MB_P2P_SCATTER", "0")
    n_reads, rl = 3_000, 150
    data = synth.reads_fastq_bytes(20_000, n_reads, rl, 0.003, 5 + k)
    fq = tmp_path / "reads.fq"
    fq.write_bytes(data)
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=12", "count", str(fq), "output", ora])
    nk, npass = multi_count(world, k, 12, data, out)
    util.assert_same_db(ora, out)
    assert sum(nk) == n_reads * (rl - k + 1)
    assert min(npass) == max(npass) and npass[0] > 1         # every rank walks the same, multi-pass plan


@pytest.mark.parametrize("world,k,wp,cap", [(8, 21, 14, 0), (4, 15, 8, 0), (2, 21, 12, 60_000)])
def test_sim_multi_rank_suffix32(tmp_path, monkeypatch, world, k, wp, cap):
    """u32 suffix keys through the NCCL route (MB_P2P_SCATTER=0): on several GPUs, too, k <= 21 travels and is sorted as u32
    suffixes below the prefix bin; routing is by bin (one message per bin and peer), every owned bin is a sort segment, count
    groups are runs of consecutive bins inside an owned file.  Same db bytes as the oracle; k-mers per rank as in the u64 form."""
    monkeypatch.setenv("MB_MULTI_SUFFIX32", "1")
    monkeypatch.setenv("MB_P2P_SCATTER", "0")
    if cap:
        monkeypatch.setenv("MB_MAX_PASS_KEYS", str(cap))
    n_reads, rl = 4_000, 150
    data = synth.reads_fastq_bytes(20_000, n_reads, rl, 0.003, 40 + k + world)
    fq = tmp_path / "reads.fq"
    fq.write_bytes(data)
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=%d" % wp, "count", str(fq), "output", ora])
    nk, npass = multi_count(world, k, wp, data, out)
    util.assert_same_db(ora, out)
    assert sum(nk) == n_reads * (rl - k + 1)
    assert min(npass) == max(npass) and (npass[0] > 1) == bool(cap)
    assert multi_key_bytes() == [4] * world


def test_sim_multi_rank_suffix32_falls_back_for_long_kmers(tmp_path, monkeypatch):
    """k = 31 does not fit 32 bits below the bin: the switch changes nothing."""
    monkeypatch.setenv("MB_MULTI_SUFFIX32", "1")
    data = synth.reads_fastq_bytes(20_000, 2_000, 150, 0.003, 3)
    fq = tmp_path / "reads.fq"
    fq.write_bytes(data)
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=31", "wprefix=12", "count", str(fq), "output", ora])
    multi_count(2, 31, 12, data, out)
    util.assert_same_db(ora, out)
    assert multi_key_bytes() == [8, 8]


@pytest.mark.parametrize("world,k,wp,cap,s32", [(4, 31, 12, 0, False), (8, 51, 10, 0, False), (2, 21, 12, 60_000, False),
                                                (8, 21, 14, 0, True), (2, 19, 12, 50_000, True)])
def test_sim_multi_rank_p2p_scatter(tmp_path, monkeypatch, world, k, wp, cap, s32):
    """The default route of round 2: extraction fused with the exchange -- every rank maps the other ranks' arenas and
    the extraction kernel stores each kept k-mer straight into the receive segment of the GPU that owns its file (in the
    emulator the ranks share one address space, so the 'mapping' is the pointer itself).  All key forms, with and without
    several passes, with and without the u32 suffix form; same db bytes as the oracle."""
    monkeypatch.setenv("MB_P2P_SCATTER", "1")
    monkeypatch.setenv("MB_MULTI_SUFFIX32", "1" if s32 else "0")
    if cap:
        monkeypatch.setenv("MB_MAX_PASS_KEYS", str(cap))
    n_reads, rl = 3_000, 150
    data = synth.reads_fastq_bytes(20_000, n_reads, rl, 0.003, 60 + k + world)
    fq = tmp_path / "reads.fq"
    fq.write_bytes(data)
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=%d" % wp, "count", str(fq), "output", ora])
    nk, npass = multi_count(world, k, wp, data, out)
    util.assert_same_db(ora, out)
    assert sum(nk) == n_reads * (rl - k + 1)
    assert (npass[0] > 1) == bool(cap)
    assert multi_key_bytes() == [4 if s32 else (8 if k <= 32 else 16)] * world


@pytest.mark.parametrize("world,k,wp,weights", [(2, 21, 12, None), (4, 31, 10, [1, 4, 2, 0]), (2, 51, 9, [5, 1]), (8, 21, 12, [3, 1, 1, 1, 2, 0, 1, 1]), (3, 21, 12, [0, 2, 5])])
def test_sim_multi_rank_count_batches(tmp_path, monkeypatch, world, k, wp, weights):
    """Count batches on several GPUs (merylOp-countThreads.C:286-380, kmers-writer-block.C:221-386): a rank whose read store
    fills closes a batch, and the count is collective, so every rank takes part in it -- the others with what they hold at
    that moment, or, once their reads are all in, from mb_count's rounds with an empty store.  Ranks with very different
    shares (and one with nothing) close different numbers of batches of their own; every rank merges the files it owns,
    rank 0 collects the histograms.  Same db as the single-process oracle."""
    monkeypatch.setenv("MB_MAX_BATCH_BASES", "60000")
    n_reads, rl = (3_000 if world == 2 and k == 21 else 1_800), 150
    data = synth.reads_fastq_bytes(20_000, n_reads, rl, 0.003, 90 + k + world)
    fq = tmp_path / "reads.fq"
    fq.write_bytes(data)
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "sim.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=%d" % wp, "count", str(fq), "output", ora])
    nk, npass = multi_count(world, k, wp, data, out, weights=weights)
    util.assert_same_db(ora, out)

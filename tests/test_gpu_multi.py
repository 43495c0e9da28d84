"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the reads are split over the ranks, k-mers
travel to the owner of their data file through the library's NCCL all-to-all, every rank writes its
own files into one directory, rank 0 the master index -- and the directory must be byte-identical to
the single-process oracle db of the whole input."""
import os

import numpy as np
import pytest
import torch

import util
import meryl_b200 as mb
from meryl_b200 import synth

pytestmark = pytest.mark.gpu

_ONLY_WORLDS = [int(x) for x in os.environ.get("MB_TEST_WORLDS", "").split(",") if x]     # e.g. MB_TEST_WORLDS=8 on an 8-GPU call (charged 8x)


def _need(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    if _ONLY_WORLDS and world not in _ONLY_WORLDS:
        pytest.skip("MB_TEST_WORLDS=%s" % os.environ["MB_TEST_WORLDS"])


def _worker(rank, world, port, fq_path, out_dir, k, wp, env=None):
    os.environ.update(env or {})
    import torch.distributed as dist
    from meryl_b200 import dist as mbd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        data = open(fq_path, "rb").read()
        recs = data.split(b"@r")[1:]                      # synthetic FASTQ: every record starts with "@r"
        first, cnt = mbd.shard_range(len(recs), rank, world)
        if os.environ.get("MB_TEST_UNEVEN"):              # very different shares: rank r gets (r * 3) % 5 parts (rank 0: nothing)
            w = [(r * 3) % 5 for r in range(world)]
            cuts = [len(recs) * sum(w[:r]) // sum(w) for r in range(world + 1)]
            first, cnt = cuts[rank], cuts[rank + 1] - cuts[rank]
        mine = b"".join(b"@r" + r for r in recs[first:first + cnt])
        # a 4 GB arena per rank: allocating the whole device and mapping seven 170 GB peer arenas costs every test of a
        # world-8 run about a minute (measured: 9 tests in 643 s); the emulator's twins run with 384 MB arenas
        c = mb.Counter(k, wp, "count", device=rank, rank=rank, world_size=world,
                       workspace_bytes=0 if os.environ.get("MB_TEST_FULL_ARENA") else (4 << 30))
        c.comm_init(mbd.broadcast_unique_id(mb.Counter.comm_unique_id, device="cuda"))
        if mine:
            c.add_file_bytes(mine)
        c.count()
        st = c.stats()
        total = mbd.all_ranks_sum(st.nKmers, "cuda")
        c.write_db(out_dir)
        c.close()
        dist.barrier()
        with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as fh:
            fh.write("%d %d %d %d %d\n" % (st.nKmers, total, st.nPasses, st.keyBytes, st.nBatches))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("k", [21, 51])
@pytest.mark.parametrize("route", ["default", "nccl-u64"])
def test_multi_gpu_db_bit_exact(tmp_path, world, k, route):
    """default: peer scatter (+ u32 suffix keys for k <= 21); nccl-u64: both switched off -- round 1's all-to-all of full keys."""
    _db_bit_exact(tmp_path, world, k, route)


def _db_bit_exact(tmp_path, world, k, route):
    _need(world)
    import torch.multiprocessing as mp
    n_reads, L = 60_000, 150
    fq = tmp_path / "reads.fq"
    fq.write_bytes(synth.reads_fastq_bytes(200_000, n_reads, L, 0.003, 17 + k))
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=12", "count", str(fq), "output", ora])
    out = str(tmp_path / "gpu.meryl")
    os.makedirs(out)
    port = 29600 + (os.getpid() + world * 7 + k + len(route)) % 300
    env = {"MB_P2P_SCATTER": "0", "MB_MULTI_SUFFIX32": "0"} if route == "nccl-u64" else None
    mp.spawn(_worker, args=(world, port, str(fq), out, k, 12, env), nprocs=world, join=True)
    stats = []
    for r in range(world):
        p = os.path.join(out, "rank%d.txt" % r)
        stats.append([int(x) for x in open(p).read().split()])
        os.remove(p)
    assert all(s[1] == n_reads * (L - k + 1) for s in stats)
    assert sum(s[0] for s in stats) == n_reads * (L - k + 1)
    util.assert_same_db(ora, out)


@pytest.mark.parametrize("world,cap", [(2, 0), (2, 700_000), (8, 0)])
def test_multi_gpu_suffix32_db_bit_exact(tmp_path, world, cap):
    """u32 suffix keys through the NCCL route (MB_P2P_SCATTER=0), routed by prefix bin (one message per bin and peer); emulator twin:
    tests/test_cusim_multi.py::test_sim_multi_rank_suffix32."""
    _need(world)
    import torch.multiprocessing as mp
    n_reads, L, k = 60_000, 150, 21
    fq = tmp_path / "reads.fq"
    fq.write_bytes(synth.reads_fastq_bytes(200_000, n_reads, L, 0.003, 91))
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=12", "count", str(fq), "output", ora])
    out = str(tmp_path / "gpu.meryl")
    os.makedirs(out)
    env = {"MB_MULTI_SUFFIX32": "1", "MB_P2P_SCATTER": "0"}
    if cap:
        env["MB_MAX_PASS_KEYS"] = str(cap)
    port = 29600 + (os.getpid() + world * 11 + cap) % 300
    mp.spawn(_worker, args=(world, port, str(fq), out, k, 12, env), nprocs=world, join=True)
    for r in range(world):
        p = os.path.join(out, "rank%d.txt" % r)
        st = [int(x) for x in open(p).read().split()]
        os.remove(p)
        assert st[1] == n_reads * (L - k + 1) and st[3] == 4 and (st[2] > 1) == bool(cap)
    util.assert_same_db(ora, out)


@pytest.mark.parametrize("world,k,cap,s32", [(2, 21, 0, False), (2, 51, 0, False), (2, 21, 700_000, False), (2, 21, 0, True), (8, 21, 0, True), (8, 31, 0, False)])
def test_multi_gpu_p2p_scatter_db_bit_exact(tmp_path, world, k, cap, s32):
    """The default route (MB_P2P_SCATTER, with and without u32 suffix keys): the extraction kernel stores every kept k-mer straight into the owning GPU's receive segment
    (peer arenas mapped with CUDA IPC, stores over NVLink); no all-to-all.  Emulator twin:
    tests/test_cusim_multi.py::test_sim_multi_rank_p2p_scatter."""
    _need(world)
    import torch.multiprocessing as mp
    n_reads, L = 60_000, 150
    fq = tmp_path / "reads.fq"
    fq.write_bytes(synth.reads_fastq_bytes(200_000, n_reads, L, 0.003, 33 + k))
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=12", "count", str(fq), "output", ora])
    out = str(tmp_path / "gpu.meryl")
    os.makedirs(out)
    env = {"MB_P2P_SCATTER": "1", "MB_MULTI_SUFFIX32": "1" if s32 else "0"}
    if cap:
        env["MB_MAX_PASS_KEYS"] = str(cap)
    port = 29600 + (os.getpid() + world * 13 + k + cap + int(s32)) % 300
    mp.spawn(_worker, args=(world, port, str(fq), out, k, 12, env), nprocs=world, join=True)
    for r in range(world):
        p = os.path.join(out, "rank%d.txt" % r)
        st = [int(x) for x in open(p).read().split()]
        os.remove(p)
        assert st[1] == n_reads * (L - k + 1) and st[3] == (4 if s32 else 8 if k <= 32 else 16) and (st[2] > 1) == bool(cap)
    util.assert_same_db(ora, out)


@pytest.mark.parametrize("world", [2, 8])
def test_cli_gpus_word(tmp_path, world):
    """`meryl-b200 gpus=N count ...`: the CLI forks one process per GPU, rank r ingests shard r of every input file
    (mb_add_file_shard), the NCCL id travels through shared memory; same directory as the oracle's."""
    _need(world)
    import subprocess
    fq, fa = tmp_path / "reads.fq", tmp_path / "contigs.fa"
    fq.write_bytes(synth.reads_fastq_bytes(200_000, 40_000, 150, 0.003, 5))
    fa.write_bytes(synth.fasta_bytes(5, 100_000, 9))
    ora, out = str(tmp_path / "ora.meryl"), str(tmp_path / "cli.meryl")
    util.run_oracle(["k=21", "memory=2", "count", str(fq), str(fa), "output", ora])
    p = subprocess.run([mb.CLI_PATH, "k=21", "memory=2", "gpus=%d" % world, "count", str(fq), str(fa), "output", out],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    util.assert_same_db(ora, out)


@pytest.mark.parametrize("world,k", [(2, 21), (2, 51), (8, 21), (8, 31)])
def test_multi_gpu_count_batches(tmp_path, world, k):
    """Count batches on several GPUs: MB_MAX_BATCH_BASES makes the ranks' read stores "fill" after 300 k bases, the shares are
    very uneven (rank 0 has no reads at all), every batch is counted by all ranks together and every rank merges the files it
    owns (engine.cu: mb_count's rounds).  Emulator twin: tests/test_cusim_multi.py::test_sim_multi_rank_count_batches."""
    _need(world)
    import torch.multiprocessing as mp
    n_reads, L = 40_000, 150
    fq = tmp_path / "reads.fq"
    fq.write_bytes(synth.reads_fastq_bytes(200_000, n_reads, L, 0.003, 71 + k))
    ora = str(tmp_path / "ora.meryl")
    util.run_oracle(["k=%d" % k, "wprefix=12", "count", str(fq), "output", ora])
    out = str(tmp_path / "gpu.meryl")
    os.makedirs(out)
    env = {"MB_MAX_BATCH_BASES": "300000", "MB_TEST_UNEVEN": "1"}
    port = 29600 + (os.getpid() + world * 17 + k) % 300
    mp.spawn(_worker, args=(world, port, str(fq), out, k, 12, env), nprocs=world, join=True)
    batches = []
    for r in range(world):
        p = os.path.join(out, "rank%d.txt" % r)
        st = [int(x) for x in open(p).read().split()]
        os.remove(p)
        assert st[1] == n_reads * (L - k + 1)
        batches.append(st[4])
    assert max(batches) >= 3, batches
    util.assert_same_db(ora, out)


@pytest.mark.parametrize("world,k,route", [(3, 21, "default"), (3, 51, "nccl-u64"), (6, 31, "default")])
def test_multi_gpu_world_not_dividing_64(tmp_path, world, k, route):
    """A world size that does not divide the 64 data files: the pass plan's last slot is a short one (files 63 / 60..63 were
    dropped before the emulator soak found it; emulator twin: test_sim_multi_rank_db_bit_exact[3|5|6|7]).  Added after round 2's
    last multi-GPU call, so it comes last."""
    _db_bit_exact(tmp_path, world, k, route)

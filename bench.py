#!/usr/bin/env python
"""bench.py -- `meryl count` throughput on B200 (BASELINE.json metric: k-mers/s end to end).

Workload (config.workload): BASELINE.json configs[1] -- k=21 canonical count of synthetic 150 bp
Illumina-shape reads (uniform-random 3 Gbp genome, both strands, 0.1 % substitutions),
`--gbp` billion bases of them PER GPU (default 100 when it fits, otherwise the largest size that
fits the device and the host, stated in config).  One step = one complete count of the workload.

  value      whole-job k-mers/s with the input resident in HBM (the 2-bit packed read store) when the
             timed region starts: prefix histogram -> extract+scatter -> radix sort -> run-length count
             -> encode, all passes, encoded db left on the device.  CUDA events on the launching stream.
  e2e        the same job through the C ABI from HOST memory: pinned host buffer of the reads' base
             stream (the reference's work-buffer format at the merylInput::loadBases seam) -> H2D ->
             pack -> count -> encoded db copied back to host images.  Wall clock around the calls.
  roofline   the radix-sort sweep kernel: algorithmic bytes N * P * 2W over its summed device time,
             against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the unmodified reference binary (oracle/_ref/meryl) on this box's host cores, timed on a
             bounded prefix of the same reads written as a FASTQ file.

`--impl reference` times only the reference binary (rank 0) and prints the same JSON shape.
Multi-GPU (torchrun, one rank per GPU): the SAME total workload is split evenly over the ranks
(strong scaling, as north_star asks: one stated base count at 1, 2, 4 and 8 GPUs); k-mers are routed to
the GPU that owns their data file by one NCCL all-to-all per pass inside the library.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 21
READ_LEN = 150
GENOME_LEN = 3_000_000_000
ERR = 0.001
SEED = 2
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "meryl")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


SYNTH_BIN = os.path.join(ROOT, "oracle", "synth_reads")
METRIC = "kmers_per_sec_meryl_count"


def cpu_reads_file(path, first, n_reads, fmt):
    """The bench's synthetic reads written by the CPU twin of the device generator (oracle/synth_reads.c, same bytes):
    the reference arm never loads CUDA or libmeryl_b200.so."""
    if not os.access(SYNTH_BIN, os.X_OK):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "synth_reads"], check=True)
    subprocess.run([SYNTH_BIN, path, str(first), str(n_reads), str(READ_LEN), str(GENOME_LEN), repr(ERR), str(SEED), str(fmt)], check=True)


def workload_name(gbp):
    return ("meryl count k=%d canonical, %.2f Gbp of synthetic %d bp reads (3 Gbp uniform-random genome, both strands, %.1f%% substitutions)"
            % (K, gbp, READ_LEN, ERR * 100))


def shared_config(total_reads, world, reduced=None):
    """`config` of the JSON line: the workload and nothing else, so that both arms print the same object for the same job."""
    gbp = total_reads * READ_LEN / 1e9
    kmers = total_reads * (READ_LEN - K + 1)
    return {"workload": workload_name(gbp) + ("; " + reduced if reduced else ""),
            "k": K, "read_len": READ_LEN, "reads": total_reads, "kmers": kmers,
            "l2": "no flush between steps: one step reads %.1f GB of packed reads and moves %.1f G k-mers through the sort, far beyond the 126 MB L2" % (total_reads * (READ_LEN + 1) * 0.375 / 1e9, kmers / 1e9),
            "parallelism": "strong scaling: the one workload split evenly over %d GPU(s)" % world}


def reference_arm(args, world, rank):
    """The reference's own CPU implementation (the unmodified binary oracle/_ref/meryl, all host threads) on a bounded
    prefix of the workload's reads, FASTQ file in /dev/shm -> database directory in /dev/shm, wall clock of the process.
    This function must not import torch, CUDA or meryl_b200: the arm is the reference alone."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if not os.access(REF_BIN, os.X_OK):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/meryl not built"}))
        return
    n_reads = int(args.ref_gbp * 1e9) // READ_LEN
    tmp = tempfile.mkdtemp(prefix="merylref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fq = os.path.join(tmp, "sample.fq")
    cpu_reads_file(fq, 0, n_reads, 1)
    n_kmers = n_reads * (READ_LEN - K + 1)
    times = []
    for it in range(args.warmup + args.steps):
        out = os.path.join(tmp, "ref.meryl")
        subprocess.run(["rm", "-rf", out])
        t0 = time.perf_counter()
        p = subprocess.run([REF_BIN, "k=%d" % K, "threads=%d" % cores, "memory=%d" % args.ref_memory_gb, "count", fq, "output", out],
                           capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            print(json.dumps({"impl": "reference", "unavailable": "reference meryl failed: " + p.stderr[-200:].replace("\n", " ")}))
            return
        if it >= args.warmup:
            times.append(dt)
    subprocess.run(["rm", "-rf", tmp])
    ms = 1000 * sum(times) / len(times)
    val = n_kmers / (ms / 1000)
    total_gbp = args.gbp if args.gbp > 0 else 100.0
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "k-mers/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": shared_config(int(total_gbp * 1e9) // READ_LEN, world),
        "sample_gbp": args.ref_gbp,
        "sample": "each step counts the first %.2f Gbp of the workload's reads (FASTQ in /dev/shm) with the unmodified reference binary" % args.ref_gbp,
        "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": cores, "kind": "reference",
                         "sample": "first %d reads (%.2f Gbp) of the workload as a FASTQ file in /dev/shm, threads=%d memory=%d, process start to exit" % (n_reads, args.ref_gbp, cores, args.ref_memory_gb)},
        "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cli_leg(gbp, local):
    """Both command line tools on the SAME FASTQ file in /dev/shm, wall clock from process start to exit: what a user
    who swaps `meryl` for `meryl-b200` sees (FASTQ parsing and the 129-file write included)."""
    cores = os.cpu_count() or 1
    n_reads = int(gbp * 1e9) // READ_LEN
    n_kmers = n_reads * (READ_LEN - K + 1)
    tmp = tempfile.mkdtemp(prefix="merylcli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    res = {"gbp": gbp, "kmers": n_kmers, "unit": "k-mers/s", "cores": cores,
           "what": "k=%d count <FASTQ in /dev/shm> output <dir in /dev/shm>, same words for both tools, wall clock of the process" % K}
    try:
        fq = os.path.join(tmp, "reads.fq")
        cpu_reads_file(fq, 0, n_reads, 1)
        res["fastq_bytes"] = os.path.getsize(fq)
        words = ["k=%d" % K, "threads=%d" % cores, "memory=64", "count", fq, "output"]
        cli = os.path.join(ROOT, "meryl_b200", "meryl-b200")
        env = dict(os.environ)
        best = None
        for it in range(2):                                           # first run warms the page cache and the driver
            subprocess.run(["rm", "-rf", os.path.join(tmp, "gpu.meryl")])
            t0 = time.perf_counter()
            p = subprocess.run([cli] + words + [os.path.join(tmp, "gpu.meryl")], capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                res["error"] = "meryl-b200: " + p.stderr[-300:]
                return res
            best = dt if best is None else min(best, dt)
        res["b200_seconds"], res["b200_value"] = best, n_kmers / best
        if os.access(REF_BIN, os.X_OK):
            t0 = time.perf_counter()
            p = subprocess.run([REF_BIN] + words + [os.path.join(tmp, "ref.meryl")], capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if p.returncode == 0:
                res["reference_seconds"], res["reference_value"] = dt, n_kmers / dt
                res["speedup"] = dt / best
                same = subprocess.run(["diff", "-rq", os.path.join(tmp, "gpu.meryl"), os.path.join(tmp, "ref.meryl")], capture_output=True, text=True)
                res["db_identical"] = (same.returncode == 0)
            else:
                res["reference_error"] = p.stderr[-300:]
    finally:
        subprocess.run(["rm", "-rf", tmp])
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gbp", type=float, default=0.0, help="billion bases of reads in total (0 = 100, reduced to what fits)")
    ap.add_argument("--ref-gbp", type=float, default=0.0, help="size of the reference's bounded sample (0 = sized so that all its steps take ~2.5 min)")
    ap.add_argument("--ref-memory-gb", type=int, default=64)
    ap.add_argument("--workspace-gb", type=float, default=0.0, help="device arena per GPU (0 = all free memory); small values keep ncu replays fast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cli-gbp", type=float, default=5.0, help="size of the FASTQ file both command line tools count (e2e_cli leg, one GPU only; 0 = skip)")
    # the other BASELINE.json configs (parity-test cases, not the headline line): e.g. --k 31, or --k 51 --read-len 15000 --err 0.005 --gbp 50
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--err", type=float, default=ERR)
    args = ap.parse_args()
    globals().update(K=args.k, READ_LEN=args.read_len, ERR=args.err)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.ref_gbp <= 0:
        # the reference binary counts 60-110 M k-mers/s on the boxes seen so far (16-32 host threads): at least 1 Gbp per
        # run (below that its start-up and teardown dominate), more when few runs are asked for; all runs within ~6 minutes
        runs = (args.steps + args.warmup) if args.impl == "reference" else 1
        args.ref_gbp = max(1.0, min(2.0, 20.0 / max(1, runs)))

    if args.impl == "reference":                    # the reference alone: no torch, no CUDA, no libmeryl_b200.so in this process
        reference_arm(args, world, rank)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import meryl_b200 as mb
    from meryl_b200 import dist as mbd
    import psutil

    free_b, total_b = torch.cuda.mem_get_info()
    host_free = psutil.virtual_memory().available

    # ---- size the workload ---------------------------------------------------------------------
    gbp = args.gbp if args.gbp > 0 else 100.0
    total_reads = int(gbp * 1e9) // READ_LEN
    rec0 = mb.synth_record_bytes(READ_LEN, 0)
    # device: the packed store (12 B per 32 positions) must leave most of HBM to the key buffers
    max_reads_dev = int(0.30 * free_b / (rec0 * 12 / 32)) * world
    # host (e2e leg): every rank pins its share of the base stream and keeps its db images
    max_reads_host = int(0.55 * host_free / rec0)
    reduced = None
    if total_reads > min(max_reads_dev, max_reads_host):
        total_reads = min(max_reads_dev, max_reads_host)
        reduced = "reduced to fit (device free %.0f GB, host free %.0f GB)" % (free_b / 1e9, host_free / 1e9)
    if world > 1:                                   # every rank must agree on the size
        t = torch.tensor([total_reads], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        total_reads = int(t.item())
    total_kmers = total_reads * (READ_LEN - K + 1)
    first_read, n_reads = mbd.shard_range(total_reads, rank, world)   # strong scaling: the same job, split

    n_est = total_reads * mb.synth_record_bytes(READ_LEN, 1)          # what `meryl` would guess: the FASTQ size
    pl = mb.plan(K, n_est, 1024 << 30)

    def new_counter(flags):
        c = mb.Counter(K, pl.w_prefix, "count", device=local, flags=flags, rank=rank, world_size=world,
                       workspace_bytes=int(args.workspace_gb * (1 << 30)))
        if world > 1:
            c.comm_init(mbd.broadcast_unique_id(mb.Counter.comm_unique_id, device="cuda"))
        return c

    chunk_reads = max(1, min(4_000_000, (600 << 20) // rec0))     # ~600 MB of base stream per generator call (15 kb reads: 40 k reads)
    dev_chunk = torch.empty(chunk_reads * rec0, dtype=torch.uint8, device="cuda")

    def ingest_device(c):
        for first in range(0, n_reads, chunk_reads):
            m = min(chunk_reads, n_reads - first)
            mb.synth_reads_device(dev_chunk.data_ptr(), first_read + first, m, READ_LEN, GENOME_LEN, ERR, SEED, 0)
            torch.cuda.synchronize()
            c.add_bases_device(dev_chunk.data_ptr(), m * rec0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident whole job ------------------------------------------------------
    sampler = ClockSampler(local)
    ms_steps, stats = [], None
    for it in range(args.warmup + args.steps):
        c = new_counter(mb.FLAG_NO_IMAGES)
        try:
            ingest_device(c)
            barrier()
            if it == args.warmup:
                sampler.start()
            c.count()
            barrier()
            st = c.stats()
            counted = mbd.all_ranks_sum(st.nKmers, "cuda") if world > 1 else st.nKmers
            assert counted == total_kmers, (counted, total_kmers)
            if it >= args.warmup:
                ms_steps.append(st.msTotal)
                stats = st
        finally:
            c.close()
    clocks = sampler.stop()
    ms = sum(ms_steps) / len(ms_steps)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = total_kmers / (ms / 1000)

    # roofline of the sort sweep kernel (rank-local)
    peaks, peak_kind = _peaks()
    sort_bytes = stats.sortLaunchKeys * 2 * stats.keyBytes
    achieved = sort_bytes / (stats.msSort / 1000) / 1e9 if stats.msSort > 0 else 0.0
    sweep_version = mb.lib().mb_dbg_sort_version(0)            # 0: query only
    sweep_name = {2: "sortx_sweep_kernel<128 threads, 8-bit digits>", 3: "sortx_sweep_kernel<256 threads, 6-bit digits>"}.get(sweep_version, "?")
    traffic = None                                              # DRAM bytes per launch, from the committed ncu capture of this kernel
    for name in ("r2_traffic_sweep.json", "r1_traffic_sweep.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                tr = json.load(fh)
            if sweep_version == 3 and stats.keyBytes == 4:
                traffic = tr["dram_bytes_per_key"] * stats.sortLaunchKeys / max(1, stats.sortPasses * stats.nPasses)
            break
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": sweep_name, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                "per": "one sweep (digit pass) over all keys of a pass over the reads; on one GPU it is issued as one launch per group of bins",
                "algorithmic_bytes_per_sweep": sort_bytes / max(1, stats.sortPasses * stats.nPasses),
                "sweeps": stats.sortPasses * stats.nPasses}

    # (The legs that start other processes come BEFORE the end-to-end leg: that one page-locks 130 GB of host memory, and a
    #  196 GB box has no room for it next to the reference binary's buckets and two databases in /dev/shm.)
    # ---- CPU baseline (rank 0, bounded sample) --------------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        sub = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                              "--ref-gbp", str(args.ref_gbp), "--ref-memory-gb", str(args.ref_memory_gb),
                              "--k", str(K), "--read-len", str(READ_LEN), "--err", str(ERR)],
                             capture_output=True, text=True, env={**os.environ, "WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0"})
        try:
            cpu = json.loads(sub.stdout.strip().splitlines()[-1]).get("cpu_baseline")
        except Exception:
            cpu = {"value": None, "unit": "k-mers/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: " + sub.stderr[-200:]}

    # ---- the command line tools side by side on one FASTQ file (rank 0, one GPU) -----------------
    cli = None
    if rank == 0 and world == 1 and args.cli_gbp > 0:
        torch.cuda.empty_cache()
        mb.lib().mb_release_device_memory()         # the arena this process keeps between counts (another process is about to use the GPU)
        try:
            cli = cli_leg(args.cli_gbp, local)
        except Exception as ex:
            cli = {"error": repr(ex)[:300]}

    # ---- e2e: host buffer -> C ABI -> host images ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty(n_reads * rec0, dtype=torch.uint8, pin_memory=True)
            for first in range(0, n_reads, chunk_reads):
                m = min(chunk_reads, n_reads - first)
                mb.synth_reads_device(dev_chunk.data_ptr(), first_read + first, m, READ_LEN, GENOME_LEN, ERR, SEED, 0)
                host[first * rec0: (first + m) * rec0].copy_(dev_chunk[: m * rec0])
            torch.cuda.synchronize()
            e_steps, est = [], None
            n_e2e = max(1, min(args.steps, 2)) + 1
            out_host = None                               # caller-owned pinned buffer for the db images (sized by the warm-up step)
            for it in range(n_e2e):
                barrier()
                t0 = time.perf_counter()
                c = new_counter(0)
                try:
                    if out_host is not None:
                        c.set_output_buffer(out_host.data_ptr(), out_host.numel())
                    c.add_bases_ptr(host.data_ptr(), n_reads * rec0, end_of_sequence=True)
                    c.count()
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    est = c.stats()
                finally:
                    c.close()
                barrier()
                if it >= 1:
                    e_steps.append(dt)
                elif out_host is None:
                    out_host = torch.empty(int(est.dbBytes * 1.02) + (64 << 20), dtype=torch.uint8, pin_memory=True)
            dt = sum(e_steps) / len(e_steps)
            if world > 1:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            e2e = {"value": total_kmers / dt, "unit": "k-mers/s", "h2d_bytes_per_step": int(est.h2dBytes),
                   "d2h_bytes_per_step": int(est.d2hBytes), "seconds_per_step": dt,
                   "input": "pinned host buffer of the base stream (1 B/base + '.' per read), output: 129 db images in (pinned) host memory"}
            del host, out_host
            if hasattr(torch._C, "_host_emptyCache"):
                torch._C._host_emptyCache()             # torch keeps freed page-locked blocks: 130 GB that the host may need back
        except Exception as ex:                 # the device-resident line above stands; say why there is no end-to-end number
            if world > 1:
                raise                           # (a rank that drops out of the collectives would hang the others)
            e2e = {"value": None, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:300]}
    del dev_chunk

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "k-mers/s",
            "value_kind": "device_resident: the whole count with the packed reads already in HBM; `e2e` is the same job from host buffers through the C ABI, `e2e_cli` the command line tools on a FASTQ file",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": shared_config(total_reads, world, reduced),
            "plan": {"w_prefix": pl.w_prefix, "key_bytes": stats.keyBytes, "passes": stats.nPasses, "sort_digit_passes": stats.sortPasses,
                     "per_rank": "%.1f GB packed read store, %.1f GB of keys per pass over the reads on rank 0" % (stats.nBases * 0.375 / 1e9, stats.nKmers * stats.keyBytes / max(1, stats.nPasses) / 1e9),
                     "routing": ("one GPU" if world == 1 else "reads split over %d GPUs; 64 data files sharded by owner(f)=min(f,63-f) mod N; the extraction kernel stores every k-mer into its owner's HBM over NVLink (peer memory), NCCL for the small collectives" % world)},
            "stage_ms": {"histogram": stats.msHistogram, "extract": stats.msExtract, "sort_hist": stats.msSortHist,
                         "sort": stats.msSort, "count": stats.msCount, "encode": stats.msEncode, "exchange": stats.msExchange},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_cli": cli, "clocks": clocks,
            "gpu_launches": int(stats.nLaunches),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

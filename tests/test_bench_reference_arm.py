"""bench.py on the CPU.  (1) The reference arm (`--impl reference`) is the reference alone: the unmodified binary oracle/_ref/meryl on reads
written by oracle/synth_reads -- the process must import neither torch nor CUDA nor the repository's shared library (round 1's
arm generated its FASTQ with the CUDA generator and so voided the driver's vs_reference ratio).  (2) The b200 arm's host logic,
dry: every leg of main() against the emulator library (tests/cusim) with torch's CUDA calls stubbed out -- it proves that the
script runs to its JSON line and that both arms describe the same job, nothing about speed."""
import json
import os
import subprocess
import sys

import pytest

import util

CODE = r"""
import sys, json
sys.argv = ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-gbp", "0.003"]
import bench
bench.main()
maps = open("/proc/self/maps").read()
print(json.dumps({"torch": "torch" in sys.modules, "meryl_b200": "meryl_b200" in sys.modules,
                  "so": [w for w in ("libmeryl_b200", "libcuda", "libcudart", "libtorch") if w in maps]}))
"""


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/meryl not built")
def test_reference_arm_is_the_reference_alone():
    p = subprocess.run([sys.executable, "-c", CODE], cwd=util.ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    line, probe = json.loads(lines[0]), json.loads(lines[1])
    assert probe == {"torch": False, "meryl_b200": False, "so": []}
    assert line["impl"] == "reference" and line["metric"] == "kmers_per_sec_meryl_count" and line["unit"] == "k-mers/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_do_nothing():
    """Under torchrun (N > 1) rank 0 alone runs the reference; the other ranks exit 0 without work and without output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], cwd=util.ROOT,
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


DRY = r"""
import sys, os
import torch
import meryl_b200 as mb
mb.LIB_PATH = sys.argv[1]
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.mem_get_info = lambda *a, **k: (8 << 30, 8 << 30)
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.empty_cache = lambda *a, **k: None
_empty = torch.empty
def empty(*a, **k):                           # emulated device memory is host memory
    k.pop("device", None); k.pop("pin_memory", None)
    return _empty(*a, **k)
torch.empty = empty
_stats = mb.Counter.stats
def stats(self):
    st = _stats(self)
    st.msTotal = st.msTotal or 1.0            # the emulator's events do not keep time
    st.msSort = st.msSort or 0.5
    return st
mb.Counter.stats = stats
sys.argv = ["bench.py", "--steps", "1", "--warmup", "1", "--gbp", "0.0003", "--workspace-gb", "0.25", "--ref-gbp", "0.003", "--cli-gbp", "0.0003"]
import bench
bench.main()
"""


def test_b200_arm_dry_run_on_the_emulator():
    sim_dir = os.path.join(util.ROOT, "tests", "cusim")
    subprocess.run(["make", "-s", "-j8", "-C", sim_dir], check=True)
    p = subprocess.run([sys.executable, "-c", DRY, os.path.join(sim_dir, "_build", "libmeryl_b200_sim.so")], cwd=util.ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "e2e_cli", "clocks", "gpu_launches"):
        assert key in line, key
    assert line["metric"] == "kmers_per_sec_meryl_count" and line["config"]["kmers"] == 2000 * 130 and line["gpu_launches"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    cli = line["e2e_cli"]                                         # the command line tool is the product: no emulator, no CPU path
    assert "no CUDA device" in cli["error"] if "error" in cli else cli["db_identical"]
    if util.have_ref():
        assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] > 0
        # the reference arm on the same job prints the same `config`
        r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--gbp", "0.0003", "--ref-gbp", "0.003"],
                           cwd=util.ROOT, capture_output=True, text=True, timeout=600)
        ref = json.loads(r.stdout.strip().splitlines()[-1])
        assert ref["config"] == line["config"] and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
        assert ref["higher_is_better"] == line["higher_is_better"] and ref["scaling"] == line["scaling"]

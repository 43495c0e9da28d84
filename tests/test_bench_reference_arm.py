"""bench.py's reference arm (`--impl reference`) is the reference alone: the unmodified binary oracle/_ref/meryl on reads
written by oracle/synth_reads -- the process must import neither torch nor CUDA nor the repository's shared library (round 1's
arm generated its FASTQ with the CUDA generator and so voided the driver's vs_reference ratio)."""
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

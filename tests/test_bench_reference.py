"""The CPU arm of bench.py (`--impl reference`): one JSON line of the contract, and ALL host cores even when the launcher pins
OMP_NUM_THREADS=1 (torchrun does: round 1's N >= 2 reference values were single-threaded while the line said cores: 32)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_uses_every_host_core_under_a_pinned_launcher():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1",
                          "--ref-ne", "4"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    cores = len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["unit"] == "DOF-updates/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["cores"] == cores and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 1e5 and line["n_gpus"] == 2 and line["steps"] == 2
    # the other ranks exit without work or output
    out1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=dict(env, RANK="1"),
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert out1.returncode == 0 and out1.stdout.strip() == ""

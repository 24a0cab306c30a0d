"""bench.py contract pieces that can be checked without a GPU: the reference arm prints one JSON line with the agreed keys
(it times the CPU restatement on a bounded sample), ranks other than 0 stay silent, and our own arm refuses to run
without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--workload", "tess-small", "--steps", "2", "--warmup", "1", "--ref-seconds", "0.5"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["metric"] == "Mrays/sec (closest-hit + shadow)" and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "tiles" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "whitted-tess-small"


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--workload", "tess-small", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    r = _run(["--workload", "tess-small", "--steps", "1", "--warmup", "3"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)

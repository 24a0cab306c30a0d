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
    r = _run(["--impl", "reference", "--workload", "tess-small", "--steps", "2", "--warmup", "1"])
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


def test_reference_arm_covers_the_sppm_half_of_the_metric():
    r = _run(["--impl", "reference", "--workload", "sppm-shadows-small", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["metric"] == "SPPM iterations/sec" and d["unit"] == "it/s" and d["value"] > 0
    assert d["config"]["workload"] == "sppm-shadows-small" and d["config"]["photons_per_iteration"] == 95 * 95
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_never_loads_the_product_library():
    """--impl reference must time the checker alone: scene description in Python, tree build and render in oracle/ -
    libtrace_cuda.so must not be mapped into that process (native_so_loaded listed it)."""
    code = (
        "import sys; sys.path.insert(0, %r); import bench\n"
        "T, o = bench.reference_setup()\n"
        "scene, camera, spp, depth = bench.build_scene(T, 'tess-small')\n"
        "osc = o.OracleScene(scene.flatten())\n"
        "s2, c2, p = bench.build_sppm_scene(T, 'sppm-caustic-glass-d5')\n"
        "s2.flatten()\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libtrace_ref' in maps and 'libtrace_cuda' not in maps, [l for l in maps.splitlines() if 'libtrace' in l]\n"
        "print('clean')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_both_arms_share_one_config():
    """The driver compares the arms' `config`: both must come from the same function of the workload alone."""
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": whitted_config(') == 2 and src.count('"config": sppm_config(') == 2

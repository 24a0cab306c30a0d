"""Debug helper (GPU box): rays on which the guarded slab test (2) and the literal one (0) disagree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, trace_jl_b200 as T
from test_gpu_parity import ray_sets
ctx = T.Context(0)
scene, camera, _ = T.scenes.shadows(resolution=256)
flat = ctx.upload(scene)
print("prims (bvh order):", flat.prims)
for name, (o, d, tmax) in ray_sets(T, scene, camera, 2_400_000, 99).items():
    ctx.set_option("slab", 0); p0, t0, _ = ctx.intersect(o, d, tmax)
    ctx.set_option("slab", 2); p2, t2, _ = ctx.intersect(o, d, tmax)
    bad = np.nonzero(p0 != p2)[0]
    print(name, "mismatches", len(bad))
    for i in bad[:8]:
        print("  o", o[i], "d", d[i], "tmax", None if tmax is None else tmax[i], "literal", p0[i], t0[i], "guarded", p2[i], t2[i])

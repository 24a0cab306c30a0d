"""Generates tests/golden/rays_<scene>.npz: seeded ray sets with the oracle's closest hit (caller-order primitive + 1,
t, barycentrics) and any-hit answers, for the three scene families of BASELINE.json (spheres + triangles, the caustic
PLY mesh with its degenerate reference tree, the tessellated generator).  The oracle had been checked against the
reference's own known answers (tests/test_oracle_known_answers.py) and published image (tests/test_golden_image.py)
when these were written; the fixtures freeze its answers so that (a) a later edit of the oracle cannot drift silently
and (b) the GPU parity tests also hold against committed vectors, not only against a freshly built checker.
    python scripts/make_golden_rays.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import trace_jl_b200 as T
import oracle_lib

N = 4096


def rays_for(flat, camera_pos, rng):
    lo, hi = flat.nodes[0]["bmin"], flat.nodes[0]["bmax"]
    ext = np.maximum(hi - lo, 1e-3)
    target = (lo + rng.random((N, 3), dtype=np.float32) * ext).astype(np.float32)
    kind = rng.integers(0, 3, size=(N, 1))
    near = (target + rng.normal(size=(N, 3)).astype(np.float32) * (0.25 * float(ext.max()))).astype(np.float32)
    inside = (lo + rng.random((N, 3), dtype=np.float32) * ext).astype(np.float32)
    origin = np.where(kind == 0, np.asarray(camera_pos, np.float32)[None], np.where(kind == 1, near, inside)).astype(np.float32)
    d = (target - origin).astype(np.float32)
    d[np.abs(d).sum(axis=1) == 0] = np.float32(1.0)
    # a few axis-parallel directions (0 * Inf = NaN slab semantics, Q21) and negative zeros (check_direction!, ray.jl:25-29)
    d[:64, 0] = 0.0
    d[64:96, 1] = -0.0
    return origin, d


def main():
    rng = np.random.default_rng(20261017)
    scenes = {
        "shadows": (T.scenes.shadows(resolution=64)[0], (0, 15, 50)),
        "caustic_glass": (T.scenes.caustic_glass()[0], (0, 150, 150)),
        "tess_small": (T.scenes.tessellated(cells=48, stacks=26, slices=24, res=(160, 90))[0], (0, 14, 15)),
    }
    for name, (scene, cam_pos) in scenes.items():
        flat = scene.flatten()
        osc = oracle_lib.OracleScene(flat)
        o, d = rays_for(flat, cam_pos, rng)
        prim, t, b = osc.intersect(o, d, slab=0)
        t_max = np.where(rng.random(N) < 0.5, np.float32(np.inf), (t * np.float32(0.999)).astype(np.float32)).astype(np.float32)
        occ = osc.occluded(o, d, t_max, slab=0)
        out = os.path.join(ROOT, "tests", "golden", f"rays_{name}.npz")
        np.savez_compressed(out, o=o, d=d, prim=prim, t=t, b=b, t_max_any=t_max, occluded=occ.astype(np.uint8),
                            n_nodes=np.int64(len(flat.nodes)), n_prims=np.int64(len(flat.prims)))
        print(name, "hits", int((prim != 0).sum()), "of", N, "occluded", int(occ.sum()), "->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

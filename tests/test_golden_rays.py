"""Committed ray fixtures (tests/golden/rays_*.npz, scripts/make_golden_rays.py): seeded ray sets over the three scene
families with the closest-hit / any-hit answers frozen.  CPU: the oracle still gives exactly these answers (no silent
drift of the checker).  GPU (-m gpu): the product, through the C ABI, gives exactly these answers - literal and guarded
box test alike - bit for bit in primitive, t and barycentrics."""
import os

import numpy as np
import pytest

import oracle_lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("shadows", "caustic_glass", "tess_small")


def _scene(T, name):
    if name == "shadows":
        return T.scenes.shadows(resolution=64)[0]
    if name == "caustic_glass":
        return T.scenes.caustic_glass()[0]
    return T.scenes.tessellated(cells=48, stacks=26, slices=24, res=(160, 90))[0]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_committed_vectors(T, name):
    g = np.load(os.path.join(GOLDEN, f"rays_{name}.npz"))
    flat = _scene(T, name).flatten()
    assert len(flat.nodes) == int(g["n_nodes"]) and len(flat.prims) == int(g["n_prims"])
    osc = oracle_lib.OracleScene(flat)
    prim, t, b = osc.intersect(g["o"], g["d"], slab=0)
    assert np.array_equal(prim, g["prim"])
    assert np.array_equal(_bits(t), _bits(g["t"])) and np.array_equal(_bits(b), _bits(g["b"]))
    assert np.array_equal(osc.occluded(g["o"], g["d"], g["t_max_any"], slab=0).astype(np.uint8), g["occluded"])
    assert (g["prim"] != 0).sum() > 2000          # the fixture is not trivially all-miss


@pytest.mark.gpu
@pytest.mark.parametrize("slab", (0, 2))
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_committed_vectors(T, ctx, name, slab):
    g = np.load(os.path.join(GOLDEN, f"rays_{name}.npz"))
    ctx.upload(_scene(T, name))
    ctx.set_option("slab", slab)
    try:
        prim, t, b = ctx.intersect(g["o"], g["d"])
        occ = ctx.occluded(g["o"], g["d"], g["t_max_any"])
    finally:
        ctx.set_option("slab", 2)
    assert np.array_equal(prim, g["prim"])
    assert np.array_equal(_bits(t), _bits(g["t"])) and np.array_equal(_bits(b), _bits(g["b"]))
    assert np.array_equal(np.asarray(occ).astype(np.uint8), g["occluded"])

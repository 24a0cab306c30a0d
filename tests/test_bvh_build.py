"""The product's host BVH builder (trace_bvh_build, csrc/bvh_build.cpp) against the oracle's literal restatement of
BVHAccel / _init / _unroll (src/accel/bvh.jl:55-206): node arrays and primitive order must be bit-identical."""
import numpy as np
import pytest

import oracle_lib


def product_build(T, bounds, max_prims=1):
    import ctypes as C
    from trace_jl_b200 import _lib as tl
    lib = tl.load()
    bounds = np.ascontiguousarray(bounds, np.float32)
    h = C.c_void_p()
    assert lib.trace_bvh_build(tl.ptr(bounds), len(bounds), max_prims, C.byref(h)) == 0
    nodes = np.zeros(lib.trace_bvh_num_nodes(h), tl.node_dtype)
    order = np.zeros(lib.trace_bvh_num_prims(h), np.uint32)
    lib.trace_bvh_copy(h, tl.ptr(nodes), tl.ptr(order))
    lib.trace_bvh_free(h)
    return nodes, order


def same(a, b):
    return a.tobytes() == b.tobytes()


def random_bounds(rng, n, clustered=False):
    c = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    if clustered:
        c = (c * np.float32(0.01) + rng.integers(-3, 4, (n, 1)).astype(np.float32)).astype(np.float32)
    e = rng.uniform(0, 0.5, (n, 3)).astype(np.float32)
    return np.concatenate([c - e, c + e], axis=1).astype(np.float32)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 1000, 20000])
@pytest.mark.parametrize("max_prims", [1, 4])
def test_builder_matches_oracle_random(T, n, max_prims):
    rng = np.random.default_rng(n * 7 + max_prims)
    for clustered in (False, True):
        b = random_bounds(rng, n, clustered)
        nodes, order = product_build(T, b, max_prims)
        rnodes, rorder, _ = oracle_lib.bvh_build(b, max_prims)
        assert len(nodes) == len(rnodes) and same(nodes, rnodes) and same(order, rorder)
        assert sorted(order.tolist()) == list(range(n))


def test_builder_degenerate_inputs(T):
    # identical centroids -> one multi-primitive leaf; empty input -> no nodes; coplanar boxes
    b = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (5, 1))
    nodes, order = product_build(T, b)
    rnodes, rorder, _ = oracle_lib.bvh_build(b)
    assert len(nodes) == 1 and same(nodes, rnodes) and same(order, rorder)
    assert (int(nodes[0]["meta"]) & 0x3FFFFFFF) == 5
    nodes, order = product_build(T, np.zeros((0, 6), np.float32))
    assert len(nodes) == 0 and len(order) == 0
    rng = np.random.default_rng(3)
    b = random_bounds(rng, 500)
    b[:, 1] = 0
    b[:, 4] = 0
    nodes, order = product_build(T, b)
    rnodes, rorder, _ = oracle_lib.bvh_build(b)
    assert same(nodes, rnodes) and same(order, rorder)


def test_preorder_invariants(T):
    rng = np.random.default_rng(11)
    nodes, order = product_build(T, random_bounds(rng, 5000))
    n = len(nodes)
    covered = np.zeros(5000, bool)
    for i, nd in enumerate(nodes):
        meta = int(nd["meta"])
        if meta >> 30 == 3:
            cnt, off = meta & 0x3FFFFFFF, int(nd["offset"])
            assert not covered[off:off + cnt].any()
            covered[off:off + cnt] = True
        else:
            assert i + 1 < n and i + 1 < int(nd["offset"]) < n
            for c in (i + 1, int(nd["offset"])):
                ch = nodes[c]
                if int(ch["meta"]) >> 30 == 3 and (int(ch["meta"]) & 0x3FFFFFFF) == 0:
                    continue            # zero-primitive leaf carries the invalid Bounds3() (Q16)
                assert np.all(ch["bmin"] >= nd["bmin"]) and np.all(ch["bmax"] <= nd["bmax"])
    assert covered.all()


def test_caustic_glass_tree_statistics(T):
    """SURVEY.md §8a1 (an independent emulation of the reference's build on docs/src/assets/models/caustic-glass.ply +
    the two floor triangles): 181 397 nodes for 88 066 primitives, 2 663 zero-primitive leaves, depth 42."""
    import os
    if not os.path.exists(T.scenes.ASSET_PLY):
        pytest.skip("asset missing")
    bvh = T.scenes._caustic_bvh(1.25)
    assert bvh.n_primitives == 88066
    rnodes, rorder, depth = oracle_lib.bvh_build(bvh.prim_bounds, 1)
    assert same(bvh.nodes, rnodes) and same(bvh.order, rorder)
    meta = bvh.nodes["meta"].astype(np.int64)
    leaves = (meta >> 30) == 3
    zero = leaves & ((meta & 0x3FFFFFFF) == 0)
    assert len(bvh.nodes) == 181397
    assert int(zero.sum()) == 2663
    assert depth == 42

"""Pins the CPU oracle (and the host-side mirror of Trace.jl's API) against every known-answer test the reference's own
suite holds for the hot path: test/test_intersection.jl, test/test_materials.jl, test/runtests.jl.
The reference compares with `≈` (rtol = sqrt(eps(Float32)) ≈ 3.45e-4) — so do these."""
import math

import numpy as np
import pytest

import oracle_lib
from oracle_lib import f3, lib, p

RTOL = math.sqrt(np.finfo(np.float32).eps)


def approx(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def scene_of(T, prims, lights=()):
    bvh = T.BVHAccel(prims, 1)
    sc = T.Scene(list(lights), bvh)
    return sc, oracle_lib.OracleScene(sc.flatten())


# ---- test/test_intersection.jl:1-20
def test_ray_bounds_intersection():
    L = lib()
    b0, b1 = f3([1, 1, 1]), f3([2, 2, 2])
    n0, n1 = f3([-2, -2, -2]), f3([-1, -1, -1])
    import ctypes as C
    t0, t1 = C.c_float(), C.c_float()
    o = f3([0, 0, 0])
    r = L.ref_bounds_intersect(p(b0), p(b1), p(o), p(f3([1, 1, 1])), np.inf, C.byref(t0), C.byref(t1))
    assert r and approx(t0.value, 1) and approx(t1.value, 2)
    r = L.ref_bounds_intersect(p(b0), p(b1), p(o), p(f3([1, 0, 0])), np.inf, C.byref(t0), C.byref(t1))
    assert not r and t0.value == 0 and t1.value == 0
    r = L.ref_bounds_intersect(p(b0), p(b1), p(f3([1.5, 1.5, 1.5])), p(f3([1, 1, 0])), np.inf, C.byref(t0), C.byref(t1))
    assert r and t0.value == 0 and approx(t1.value, 0.5)
    for slab in (0, 1):
        assert L.ref_bounds_intersect_p(p(b0), p(b1), p(o), p(f3([1, 1, 1])), np.inf, slab)
        assert not L.ref_bounds_intersect_p(p(n0), p(n1), p(o), p(f3([1, 1, 1])), np.inf, slab)


# ---- test/test_intersection.jl:22-87
def test_ray_sphere_intersection(T):
    core = T.ShapeCore(T.Transformation(), False)
    s = T.Sphere(core, 1.0, 360.0)
    sc, osc = scene_of(T, [T.GeometricPrimitive(s)])
    for o, d, pt, n in (([0, -2, 0], [0, 1, 0], [0, -1, 0], [0, -1, 0]), ([0, 0, -2], [0, 0, 1], [0, 0, -1], [0, 0, -1])):
        prim, t, _ = osc.intersect([o], [d])
        assert prim[0] == 1 and approx(t[0], 1.0)
        assert osc.occluded([o], [d])[0]
        hr = osc.hit_record(o, d)
        assert approx(np.array(o) + np.array(d) * t[0], pt)
        assert approx(hr["p"], pt) and approx(hr["ng"], n)
        assert approx(np.linalg.norm(hr["ng"]), 1) and approx(np.linalg.norm(hr["ns"]), 1)
    # spawned ray leaving the surface misses (spawn_ray: o = p + 1e-6 * d)
    hr = osc.hit_record([0, -2, 0], [0, 1, 0])
    d = np.array([0, -1, 0], np.float32)
    o = hr["p"] + np.float32(1e-6) * d
    assert approx(o, hr["p"])
    prim, _, _ = osc.intersect([o], [d])
    assert prim[0] == 0
    # inside the sphere
    prim, t, _ = osc.intersect([[0, 0, 0]], [[0, 1, 0]])
    hr = osc.hit_record([0, 0, 0], [0, 1, 0])
    assert prim[0] == 1 and approx(t[0], 1) and approx(hr["ng"], [0, 1, 0])
    # origin on the surface
    prim, t, _ = osc.intersect([[0, -1, 0]], [[0, -1, 0]])
    hr = osc.hit_record([0, -1, 0], [0, -1, 0])
    assert prim[0] == 1 and abs(t[0]) < 1e-6 and approx(hr["p"], [0, -1, 0]) and approx(hr["ng"], [0, -1, 0])
    # translated sphere
    core = T.ShapeCore(T.translate([0, 2, 0]), False)
    sc, osc = scene_of(T, [T.GeometricPrimitive(T.Sphere(core, 1.0, 360.0))])
    prim, t, _ = osc.intersect([[0, 0, 0]], [[0, 1, 0]])
    hr = osc.hit_record([0, 0, 0], [0, 1, 0])
    assert prim[0] == 1 and approx(t[0], 1) and approx(hr["p"], [0, 1, 0]) and approx(hr["ng"], [0, -1, 0])
    assert osc.occluded([[0, 0, 0]], [[0, 1, 0]])[0]


# ---- test/test_intersection.jl:89-127
def test_triangle(T):
    core = T.ShapeCore(T.translate([0, 0, 2]), False)
    tris = T.create_triangle_mesh(core, 1, [1, 2, 3], 3, [[0, 0, 0], [1, 0, 0], [1, 1, 0]], [[0, 0, -1]] * 3)
    tv = tris[0].vertices()
    assert approx(tris[0].area(), np.linalg.norm(tv[0] - tv[1]) ** 2 * 0.5)
    assert tris[0].world_bound().approx(T.Bounds3([0, 0, 2], [1, 1, 2]))
    assert tris[0].object_bound().approx(T.Bounds3([0, 0, 0], [1, 1, 0]))
    sc, osc = scene_of(T, [T.GeometricPrimitive(tris[0])])
    # the reference calls intersect(triangle, ray) directly (axis-parallel rays through a vertex / an edge)
    hr = osc.hit_record([0, 0, -2], [0, 0, 1], prim=0)
    assert hr is not None
    assert approx(hr["t"], 4) and approx(hr["p"], [0, 0, 2]) and np.allclose(hr["uv"], [0, 0], atol=1e-6)
    assert approx(hr["ng"], [0, 0, -1]) and approx(hr["wo"], [0, 0, -1])
    hr = osc.hit_record([1, 0.5, 0], [0, 0, 1], prim=0)
    assert hr is not None
    assert approx(hr["t"], 2) and approx(hr["p"], [1, 0.5, 2]) and approx(hr["uv"], [1, 0.5])
    assert approx(hr["ng"], [0, 0, -1]) and approx(hr["wo"], [0, 0, -1])
    # through the BVH the same rays are culled: d.x = d.y = 0 gives 0 * Inf = NaN in the slab test and
    # `tx_min < t_max` is false for NaN (bounds.jl:180-200)
    prim, _, _ = osc.intersect([[0, 0, -2]], [[0, 0, 1]])
    assert prim[0] == 0
    prim, t, _ = osc.intersect([[0.75, 0.25, -2]], [[1e-3, 1e-3, 1]])
    assert prim[0] == 1 and approx(t[0], 4)


# ---- test/test_intersection.jl:129-156 (BVH nested inside a BVH)
def test_bvh_nested(T):
    prims = []
    for i in range(0, 22, 3):
        prims.append(T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([i, i, 0]), False), 1.0, 360.0)))
    bvh = T.BVHAccel(prims[:4])
    bvh2 = T.BVHAccel(prims[4:] + [bvh])
    assert bvh.world_bound().approx(T.Bounds3([-1, -1, -1], [10, 10, 1]))
    assert bvh2.world_bound().approx(T.Bounds3([-1, -1, -1], [22, 22, 1]))
    osc = oracle_lib.OracleScene(T.Scene([], bvh2).flatten())
    prim, t, _ = osc.intersect([[-2, 0, 0]], [[1, 0, 0]])
    hr = osc.hit_record([-2, 0, 0], [1, 0, 0])
    assert prim[0] != 0 and approx(t[0], 1) and approx(hr["p"], [-1, 0, 0])
    prim, t, _ = osc.intersect([[0, 18, 0]], [[1, 0, 0]])
    hr = osc.hit_record([0, 18, 0], [1, 0, 0])
    assert prim[0] != 0 and approx(t[0], 17) and approx(hr["p"], [17, 18, 0])


# ---- test/test_intersection.jl:158-195
def test_bvh_spheres_in_a_row(T):
    prims = [T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.Transformation(), False), 1.0, 360.0)),
             T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0, 0, 4]), False), 2.0, 360.0)),
             T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0, 0, 11]), False), 4.0, 360.0))]
    bvh = T.BVHAccel(prims)
    assert bvh.world_bound().approx(T.Bounds3([-4, -4, -1], [4, 4, 15]))
    osc = oracle_lib.OracleScene(T.Scene([], bvh).flatten())
    _, t, _ = osc.intersect([[0, 0, -2]], [[0, 0, 1]])
    assert approx(t[0], 1)
    _, t, _ = osc.intersect([[1.5, 0, -2]], [[0, 0, 1]])
    hr = osc.hit_record([1.5, 0, -2], [0, 0, 1])
    assert 2 < t[0] < 6 and approx(np.array([1.5, 0, -2]) + np.array([0, 0, 1]) * t[0], hr["p"])
    _, t, _ = osc.intersect([[3, 0, -2]], [[0, 0, 1]])
    assert 7 < t[0] < 15


# ---- test/test_materials.jl
def test_fresnel_dielectric():
    assert abs(lib().ref_fresnel_dielectric(1.0, 1.0, 1.0)) < 1e-7
    assert abs(lib().ref_fresnel_dielectric(0.5, 1.0, 1.0)) < 1e-7


def test_fresnel_specular_normal_incidence():
    out = np.zeros(8, np.float32)
    one = f3([1, 1, 1])
    lib().ref_fresnel_specular_sample(p(one), p(one), 1.0, 1.0, p(f3([0, 0, 1])), p(f3([0, 0])), p(out))
    assert approx(out[:3], [0, 0, -1]) and approx(out[3], 1.0)
    assert int(out[7]) == (16 | 2)          # BSDF_SPECULAR | BSDF_TRANSMISSION


def test_microfacet_reflection_normal_incidence():
    out = np.zeros(8, np.float32)
    one = f3([1, 1, 1])
    lib().ref_microfacet_reflection_sample(p(one), 1.0, 1.0, 0, 1.0, 1.0, p(f3([0, 0, 1])), p(f3([0, 0])), p(out))
    assert approx(out[:3], [0, 0, 1])


def test_microfacet_transmission_normal_incidence():
    """test/test_materials.jl:56-68: TrowbridgeReitz(1, 1), eta 1 -> 2, wo = +z, u = (0, 0)  =>  wi = (0, 0, -1)."""
    out = np.zeros(8, np.float32)
    one = f3([1, 1, 1])
    lib().ref_microfacet_transmission_sample(p(one), 1.0, 1.0, 1.0, 2.0, p(f3([0, 0, 1])), p(f3([0, 0])), p(out))
    assert approx(out[:3], [0, 0, -1])
    assert int(out[7]) == -1                # the reference's sample_f returns `nothing` as the sampled type (:319)


def test_bxdf_type_flags(T):
    import ctypes as C
    # SpecularReflection & (SPECULAR|REFLECTION); SpecularTransmission & (SPECULAR|TRANSMISSION); FresnelSpecular & all three
    from trace_jl_b200 import _lib as tl
    m = np.zeros(1, tl.material_dtype)
    m[0] = (tl.MAT_MIRROR, [1, 1, 1], [0, 0, 0], 1.0, 0.0, 0.0, 0)
    out = np.zeros(8, np.float32)
    n = lib().ref_bsdf_sample(p(m), 0, p(f3([0, 0, 1])), p(f3([0.5, 0.5])), 16 | 1, p(out))
    assert n == 1 and int(out[7]) == (16 | 1) and approx(out[:3], [0, 0, 1])
    m[0] = (tl.MAT_GLASS, [1, 1, 1], [1, 1, 1], 1.5, 0.0, 0.0, 1)
    assert lib().ref_bsdf_sample(p(m), 0, p(f3([0, 0, 1])), p(f3([0.5, 0.5])), 16 | 2, p(out)) == 2
    assert int(out[7]) == (16 | 2) and approx(out[:3], [0, 0, -1])
    assert lib().ref_bsdf_sample(p(m), 1, p(f3([0, 0, 1])), p(f3([0.5, 0.5])), 31, p(out)) == 1
    assert int(out[7]) == (16 | 2)


# ---- test/runtests.jl:11-32
def test_bounds2_iteration(T):
    b = T.Bounds2([1, 3], [4, 4])
    targets = [(1, 3), (2, 3), (3, 3), (4, 3), (1, 4), (2, 4), (3, 4), (4, 4)]
    assert len(b) == 8
    assert [tuple(q) for q in b] == targets
    b = T.Bounds2([-1, -1], [1, 1])
    assert len(b) == 9
    assert [tuple(q) for q in b] == [(-1, -1), (0, -1), (1, -1), (-1, 0), (0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]


# ---- test/runtests.jl:34-41
def test_sphere_bound(T):
    s = T.Sphere(T.ShapeCore(T.translate([0, 0, 0]), False), 1.0, -1.0, 1.0, 360.0)
    sb = s.object_bound()
    assert np.all(sb.p_min == -1) and np.all(sb.p_max == 1)


# ---- test/runtests.jl:43-48
def test_lanczos_filter(T):
    l = T.LanczosSincFilter([4, 4], 3.0)
    assert approx(l([0, 0]), 1.0) and l([4, 4]) < 1e-6 and abs(l([5, 5])) < 1e-7
    for q in ([0.3, 1.7], [2.5, 0.1], [3.9, 3.9], [0, 0]):
        assert abs(float(l(q)) - lib().ref_lanczos(q[0], q[1], 4.0, 4.0, 3.0)) < 1e-6


# ---- test/runtests.jl:50-58
def test_film(T):
    film = T.Film([1920, 1080], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([4, 4], 3.0), 35.0, 1.0, None)
    assert film.pixels.shape[:2] == (1080, 1920)
    assert film.get_sample_bounds() == T.Bounds2([-3, -3], [1924, 1084])


# ---- test/runtests.jl:60-133 (FilmTile bounds, add_sample!, merge)
def _tile_bounds(film, smin, smax):
    r = film.filter.radius
    p0 = np.ceil(np.array(smin, np.float32) - 0.5 - r)
    p1 = np.floor(np.array(smax, np.float32) - 0.5 + r) + 1
    lo = np.maximum(p0, film.crop_bounds.p_min)
    hi = np.minimum(p1, film.crop_bounds.p_max)
    return [int(lo[0]), int(lo[1]), int(hi[0]), int(hi[1])]


def test_film_tile(T):
    import ctypes as C
    film = T.Film([1920, 1080], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([4, 4], 3.0), 35.0, 1.0, None)
    fd = film.desc()
    tb = _tile_bounds(film, [1, 1], [10, 10])
    assert tb == [1, 1, 14, 14]
    w = np.zeros((14, 14), np.float32)
    c = np.zeros((14, 14, 3), np.float32)
    tbv = np.array(tb, np.int32)
    lib().ref_film_tile_add_sample(C.byref(fd), p(tbv), 1.0, 1.0, p(f3([1, 1, 1])), p(c), p(w))
    for i, j in zip(range(0, 4), range(1, 5)):
        assert w[i, i] > 0 and w[j, j] > 0 and w[i, i] > w[j, j]
    tb = _tile_bounds(film, [10, 10], [60, 60])
    assert tb == [6, 6, 64, 64]
    w = np.zeros((59, 59), np.float32)
    c = np.zeros((59, 59, 3), np.float32)
    tbv = np.array(tb, np.int32)
    lib().ref_film_tile_add_sample(C.byref(fd), p(tbv), 20.0, 20.0, p(f3([1, 1, 1])), p(c), p(w))
    d = np.diag(w)
    for i, j in zip(range(10, 14), range(17, 13, -1)):          # symmetrical (1-based 11:14 vs 18:-1:15)
        assert approx(d[i], d[j])
    for i, j in zip(range(10, 13), range(11, 14)):              # increasing left to right
        assert 0 < d[i] < d[j]
    for i, j in zip(range(15, 18), range(16, 19)):              # decreasing
        assert d[i] > d[j] > 0
    assert np.allclose(c[..., 0], w)


# ---- test/runtests.jl:135-170
def test_perspective_camera(T):
    film = T.Film([1920, 1080], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([4, 4], 3.0), 35.0, 1.0, None)
    cam = T.PerspectiveCamera(T.translate([0, 0, 0]), T.Bounds2([0, 0], [10, 10]), 0.0, 1.0, 0.0, 700.0, 45.0, film)

    def gen(fx, fy):
        pc = cam.raster_to_camera.point([fx, fy, 0])
        d = T.normalize(pc)
        o = cam.camera_to_world.point([0, 0, 0])
        return o, T.normalize(cam.camera_to_world.vector(d))

    o1, d1 = gen(1, 1)
    o2, d2 = gen(1920, 1920)
    assert np.all(o1 == 0) and np.all(o2 == 0)
    assert d1[0] < d2[0] and d1[1] < d2[1]
    assert np.argmax(np.abs(d1)) == 2 and np.argmax(np.abs(d2)) == 2
    _, dx = gen(2, 1)
    _, dy = gen(1, 2)
    assert dx[0] > d1[0] and approx(dx[1], d1[1])
    assert approx(dy[0], d1[0]) and dy[1] > d1[1]


# ---- sampler/sampling.jl:43-76 (no reference test exists; van der Corput / Halton closed forms)
def test_radical_inverse():
    L = lib()
    assert L.ref_radical_inverse(0, 1) == 0.5 and L.ref_radical_inverse(0, 2) == 0.25 and L.ref_radical_inverse(0, 3) == 0.75
    assert abs(L.ref_radical_inverse(1, 1) - 1 / 3) < 1e-7 and abs(L.ref_radical_inverse(1, 5) - (2 / 3 + 1 / 9)) < 1e-6
    assert abs(L.ref_radical_inverse(2, 7) - (2 / 5 + 1 / 25)) < 1e-6
    assert L.ref_radical_inverse(3, 0) == 0.0


def test_loose_slab_quirk_q26_work_counts(T):
    """the reference's box test keeps the LARGER y far bound (bounds.jl:191), which makes it
    accept far more boxes than a textbook slab test without changing any hit.  The survey's independent emulation of
    the literal build + traversal on C1 measured ~406-520 box tests and ~48 triangle tests per ray with the literal
    test against ~40 / ~2 with the textbook one (7 396 camera rays mixed with random rays), and 0 hit mismatches.
    The restatement shows the same signature on the 86 x 86 camera-ray grid of C1: a restatement that had silently
    "fixed" line 191 would sit at ~20 box tests per ray."""
    scene, camera, _ = T.scenes.caustic_glass()
    osc = __import__("oracle_lib").OracleScene(scene.flatten())
    xs = np.arange(1, 257, 3, dtype=np.float32) + 0.5
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    pts = camera.raster_to_camera.points(np.stack([X.ravel(), Y.ravel(), np.zeros(X.size, np.float32)], 1))
    d = pts / np.linalg.norm(pts, axis=1, keepdims=True)
    d = (d @ camera.camera_to_world.m[:3, :3].T).astype(np.float32)
    o = np.tile(camera.camera_to_world.point([0, 0, 0])[None], (len(d), 1)).astype(np.float32)
    assert len(o) == 7396
    lit = osc.intersect(o, d, slab=0, counters=True)
    std = osc.intersect(o, d, slab=1, counters=True)
    grd = osc.intersect(o, d, slab=2, counters=True)
    n = float(len(o))
    assert 400 < lit[3][0] / n < 700 and 40 < lit[3][1] / n < 70          # measured 559.5 / 53.7
    assert 10 < std[3][0] / n < 45 and 1.5 < std[3][1] / n < 3.0          # measured 19.8 / 2.17
    assert grd[3][0] <= std[3][0] * 1.05                                  # the guarded test does the textbook amount of work
    assert 20 <= int(lit[3][2]) < 64                                      # deepest pending-node stack (survey: 27 seen; limit 64, bvh.jl:222)
    for other in (std, grd):                                              # ... and none of them changes a hit on this set
        assert np.array_equal(lit[0], other[0]) and np.array_equal(lit[1].view(np.uint32), other[1].view(np.uint32))

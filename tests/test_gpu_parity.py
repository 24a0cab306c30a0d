"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): closest-hit primitive ids bit-exact, t within 2 ULP (we observe and
assert bit-equality of t and barycentrics on these sets), any-hit booleans exact; images within a stated tolerance."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
MISMATCH_LOG = []


# ---------------------------------------------------------------- helpers
def ulp_diff(a, b):
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def camera_rays(T, camera, n, rng):
    film = camera.film
    sb = film.get_sample_bounds()
    fx = rng.uniform(sb.p_min[0], sb.p_max[0] + 1, n).astype(np.float32)
    fy = rng.uniform(sb.p_min[1], sb.p_max[1] + 1, n).astype(np.float32)
    pts = camera.raster_to_camera.points(np.stack([fx, fy, np.zeros(n, np.float32)], 1))
    d = pts / np.linalg.norm(pts, axis=1, keepdims=True).astype(np.float32)
    m = camera.camera_to_world.m[:3, :3]
    d = (d @ m.T).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = np.tile(camera.camera_to_world.point([0, 0, 0])[None], (n, 1)).astype(np.float32)
    return o, d


def bbox_of(flat):
    lo = np.array(flat.nodes[0]["bmin"], np.float32)
    hi = np.array(flat.nodes[0]["bmax"], np.float32)
    return lo, hi


def ray_sets(T, scene, camera, n, seed):
    rng = np.random.default_rng(seed)
    flat = scene.flatten()
    lo, hi = bbox_of(flat)
    c, rad = (lo + hi) / 2, float(np.linalg.norm(hi - lo)) / 2
    sets = {}
    sets["R1_camera"] = camera_rays(T, camera, n, rng) + (None,)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (c + 1.5 * rad * v).astype(np.float32)
    tgt = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    sets["R2_sphere_to_box"] = (o, (tgt - o).astype(np.float32), None)
    # R3 secondary rays from points inside the box in random directions, finite t_max on half of them
    o3 = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d3 = rng.normal(size=(n, 3)).astype(np.float32)
    t3 = np.where(rng.uniform(size=n) < 0.5, np.float32(np.inf), rng.uniform(0.01, 2 * rad, n)).astype(np.float32)
    sets["R3_interior"] = (o3, d3, t3)
    # R4 shadow rays p -> light exactly as spawn_ray(p0, p1) builds them (un-normalised d, t_max = Inf)
    if len(flat.lights):
        lp = np.array(flat.lights[0]["position"], np.float32)
        p0 = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
        dd = (lp[None] - p0).astype(np.float32)
        sets["R4_shadow"] = ((p0 + np.float32(1e-6) * dd).astype(np.float32), dd, None)
    # R5 adversarial: through triangle vertices / edge midpoints, axis-parallel, zero and negative-zero components
    if len(flat.tri_vertices):
        k = rng.integers(0, len(flat.tri_vertices), n)
        tv = flat.tri_vertices[k]
        w = rng.integers(0, 3, n)
        vert = tv[np.arange(n), w]
        mid = ((tv[np.arange(n), w] + tv[np.arange(n), (w + 1) % 3]) * np.float32(0.5)).astype(np.float32)
        tgt = np.where((rng.uniform(size=n) < 0.5)[:, None], vert, mid).astype(np.float32)
        o5 = (c + 1.2 * rad * v).astype(np.float32)
        d5 = (tgt - o5).astype(np.float32)
        ax = rng.integers(0, 3, n)
        sel = rng.uniform(size=n) < 0.3
        for a in range(3):                                  # axis-parallel rays aimed at the target
            m = sel & (ax == a)
            o5[m] = tgt[m]
            o5[m, a] = (c[a] - 1.5 * rad)
            d5[m] = 0
            d5[m, a] = 1
        negz = rng.uniform(size=n) < 0.15
        d5[negz & (d5[:, 0] == 0), 0] = np.float32(-0.0)
        sets["R5_adversarial"] = (o5, d5, None)
    return sets


def check_scene(T, ctx, scene, camera, n, seed, label):
    flat = ctx.upload(scene)
    osc = oracle_lib.OracleScene(flat)
    summary = []
    for name, (o, d, tmax) in ray_sets(T, scene, camera, n, seed).items():
        rprim, rt, rb = osc.intersect(o, d, tmax, slab=0)
        rocc = osc.occluded(o, d, tmax, slab=0)
        ctx.set_option("walk", 0)               # the reference loop (one node per step) first
        for slab in (0, 1, 2):
            ctx.set_option("slab", slab)
            prim, t, b = ctx.intersect(o, d, tmax)
            occ = ctx.occluded(o, d, tmax)
            bad_prim = int(np.count_nonzero(prim != rprim))
            hit = (rprim != 0) & (prim == rprim)
            ulps = ulp_diff(t[hit], rt[hit])
            summary.append((label, name, slab, len(o), int(hit.sum()), bad_prim, int(ulps.max()) if len(ulps) else 0))
            if slab == 1:
                # the textbook slab test (no robustness margin) is NOT hit-equivalent to the reference's test on
                # zero-thickness boxes: it is measured and reported, never asserted, and never the default
                MISMATCH_LOG.append((label, name, len(o), bad_prim, int(np.count_nonzero(occ != rocc))))
                continue
            assert bad_prim == 0, f"{label}/{name}/slab{slab}: {bad_prim} closest-hit primitive ids differ"
            assert np.array_equal(t[hit].view(np.uint32), rt[hit].view(np.uint32)), f"{label}/{name}/slab{slab}: t not bit-equal (max {ulps.max()} ULP)"
            assert np.array_equal(t[~hit].view(np.uint32), rt[~hit].view(np.uint32))
            assert np.array_equal(b.view(np.uint32), rb.view(np.uint32)), f"{label}/{name}/slab{slab}: barycentrics differ"
            assert np.array_equal(occ, rocc), f"{label}/{name}/slab{slab}: {np.count_nonzero(occ != rocc)} any-hit results differ"
            # consistency: with t_max = Inf every closest hit is also an any-hit
            if tmax is None:
                assert np.array_equal(occ, prim != 0)
        # the warp-synchronous loop with batched leaves (option "leaf_wait") changes WHEN a lane tests a leaf's primitives,
        # never which nodes / primitives a ray visits or in which order: same bits as the oracle
        for slab, lw in ((2, 8), (2, 4), (0, 16), (2, 32), (2, 16), (0, 8)):
            ctx.set_option("slab", slab)
            ctx.set_option("leaf_wait", lw)
            prim, t, b = ctx.intersect(o, d, tmax)
            occ = ctx.occluded(o, d, tmax)
            assert np.array_equal(prim, rprim), f"{label}/{name}/slab{slab}/leaf_wait{lw}: primitive ids differ"
            assert np.array_equal(t.view(np.uint32), rt.view(np.uint32)), f"{label}/{name}/slab{slab}/leaf_wait{lw}: t differs"
            assert np.array_equal(b.view(np.uint32), rb.view(np.uint32)), f"{label}/{name}/slab{slab}/leaf_wait{lw}: barycentrics differ"
            assert np.array_equal(occ, rocc), f"{label}/{name}/slab{slab}/leaf_wait{lw}: any-hit differs"
        ctx.set_option("leaf_wait", 0)
        # the pair-node walk (option "walk" = 1: both children's boxes in the parent, far child pushed with its entry
        # distance) visits the same primitives in the same order with the same t_max updates: same bits as the oracle
        for slab in (0, 2):
            ctx.set_option("slab", slab)
            ctx.set_option("walk", 1)
            prim, t, b = ctx.intersect(o, d, tmax)
            occ = ctx.occluded(o, d, tmax)
            assert np.array_equal(prim, rprim), f"{label}/{name}/slab{slab}/pair: {np.count_nonzero(prim != rprim)} primitive ids differ"
            assert np.array_equal(t.view(np.uint32), rt.view(np.uint32)), f"{label}/{name}/slab{slab}/pair: t differs"
            assert np.array_equal(b.view(np.uint32), rb.view(np.uint32)), f"{label}/{name}/slab{slab}/pair: barycentrics differ"
            assert np.array_equal(occ, rocc), f"{label}/{name}/slab{slab}/pair: any-hit differs"
        ctx.set_option("slab", 2)               # back to the defaults: guarded box test, pair-node walk
    for s in summary:
        print("parity", *s)
    return summary


# ---------------------------------------------------------------- ray-query parity
def test_reference_test_scenes(T, ctx):
    """The scenes of test/test_intersection.jl:129-195 (8 spheres with a BVH nested in a BVH; 3 spheres in a row)."""
    prims = []
    for i in range(0, 22, 3):
        prims.append(T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([i, i, 0]), False), 1.0, 360.0)))
    bvh2 = T.BVHAccel(prims[4:] + [T.BVHAccel(prims[:4])])
    scene = T.Scene([], bvh2)
    flat = ctx.upload(scene)
    prim, t, _ = ctx.intersect([[-2, 0, 0], [0, 18, 0]], [[1, 0, 0], [1, 0, 0]])
    assert prim[0] != 0 and prim[1] != 0 and abs(t[0] - 1) < 1e-5 and abs(t[1] - 17) < 1e-4
    film = T.Film([64, 64], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    cam = T.PerspectiveCamera(T.look_at([10, 10, 60], [10, 10, 0]), T.Bounds2([-3000, 3000], [3000, 9000]), 0, 1, 0, 1e6, 90.0, film)
    check_scene(T, ctx, scene, cam, 20000, 1, "nested-bvh")
    prims = [T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.Transformation(), False), 1.0, 360.0)),
             T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0, 0, 4]), False), 2.0, 360.0)),
             T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0, 0, 11]), False), 4.0, 360.0))]
    scene = T.Scene([], T.BVHAccel(prims))
    ctx.upload(scene)
    prim, t, _ = ctx.intersect([[0, 0, -2], [1.5, 0, -2], [3, 0, -2]], [[0, 0, 1]] * 3)
    assert abs(t[0] - 1) < 1e-6 and 2 < t[1] < 6 and 7 < t[2] < 15
    check_scene(T, ctx, scene, cam, 20000, 2, "spheres-row")


def test_shadows_scene_rays(T, ctx):
    scene, camera, _ = T.scenes.shadows(resolution=256)
    check_scene(T, ctx, scene, camera, 100000, 3, "shadows")


def test_partial_spheres_and_flipped_shapes(T, ctx):
    """Clipped spheres (z_min / z_max / phi_max), reverse_orientation, a scaled (handedness-swapping) transform."""
    mat = T.MatteMaterial(T.ConstantTexture(T.RGBSpectrum(0.5)), T.ConstantTexture(0.0))
    prims = [T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0, 0, 0]), False), 1.0, -0.5, 0.7, 300.0), mat),
             T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([3, 0, 0]), True), 1.2, 360.0), mat),
             T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0, 3, 0]) * T.scale(1, -1, 1), False), 0.8, -0.8, 0.3, 200.0), mat)]
    tris = T.create_triangle_mesh(T.ShapeCore(T.translate([-3, 0, 0]), True), 2, [1, 2, 3, 1, 3, 4], 4,
                                  [[0, 0, 0], [1, 0, 0], [1, 1, 0.5], [0, 1, 0]])
    prims += [T.GeometricPrimitive(t, mat) for t in tris]
    scene = T.Scene([T.PointLight(T.translate([0, 5, 5]), T.RGBSpectrum(10.0))], T.BVHAccel(prims))
    film = T.Film([64, 64], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    cam = T.PerspectiveCamera(T.look_at([0, 1, 12], [0, 1, 0]), T.Bounds2([-3000, -3000], [3000, 3000]), 0, 1, 0, 1e6, 90.0, film)
    check_scene(T, ctx, scene, cam, 50000, 4, "partial-spheres")


def test_random_triangle_soup(T, ctx):
    rng = np.random.default_rng(5)
    n = 20000
    c = rng.uniform(-5, 5, (n, 1, 3))
    v = (c + rng.normal(scale=0.15, size=(n, 3, 3))).astype(np.float32)
    v[:50, 2] = v[:50, 1]                                     # degenerate triangles (is_degenerate)
    verts = v.reshape(-1, 3)
    idx = np.arange(1, 3 * n + 1, dtype=np.uint32)
    ident = T.ShapeCore(T.Transformation(), False)
    mat = T.MatteMaterial(T.ConstantTexture(T.RGBSpectrum(0.5)), T.ConstantTexture(0.0))
    ts = T.TriangleSet(ident, T.TriangleMesh(ident.object_to_world, n, idx, 3 * n, verts))
    scene = T.Scene([T.PointLight(T.translate([0, 9, 0]), T.RGBSpectrum(10.0))], T.BVHAccel([T.PrimitiveBatch(ts, mat)], 1))
    film = T.Film([64, 64], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    cam = T.PerspectiveCamera(T.look_at([0, 0, 20], [0, 0, 0]), T.Bounds2([-3000, -3000], [3000, 3000]), 0, 1, 0, 1e6, 90.0, film)
    check_scene(T, ctx, scene, cam, 100000, 6, "soup")
    # leaves with several primitives (max_node_primitives = 4): later primitive wins ties, same order as the oracle
    scene4 = T.Scene(scene.lights, T.BVHAccel([T.PrimitiveBatch(ts, mat)], 4))
    check_scene(T, ctx, scene4, cam, 50000, 7, "soup-leaf4")


def test_caustic_glass_rays(T, ctx):
    if not os.path.exists(T.scenes.ASSET_PLY):
        pytest.skip("asset missing")
    scene, camera, _ = T.scenes.caustic_glass(resolution=256)
    check_scene(T, ctx, scene, camera, 100000, 8, "caustic-glass")


def test_empty_and_tiny_inputs(T, ctx):
    scene, camera, _ = T.scenes.shadows(resolution=32)
    ctx.upload(scene)
    prim, t, b = ctx.intersect(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(prim) == 0 and len(t) == 0
    prim, t, b = ctx.intersect([[0.5, 0.5, 5]], [[0, 0, -1]])
    assert prim.shape == (1,)
    # NaN / Inf rays must not hang or crash and must agree with the oracle
    o = np.array([[0.5, 0.5, 5], [np.nan, 0, 0], [0.5, 0.5, 5], [0.5, 0.5, 5]], np.float32)
    d = np.array([[0, 0, 0], [0, 0, -1], [np.inf, 0, -1], [np.nan, 1, 1]], np.float32)
    osc = oracle_lib.OracleScene(scene.flatten())
    prim, t, _ = ctx.intersect(o, d)
    rprim, rt, _ = osc.intersect(o, d)
    assert np.array_equal(prim, rprim)


def test_guarded_slab_equals_literal_at_scale(T, ctx):
    """before a cheaper box test may be the default it has to show ZERO closest-hit / any-hit
    mismatches against the reference's literal test on >= 10^7 rays per scene.  Literal-on-GPU is itself pinned to the
    oracle above; here both variants run on the GPU over 1.2e7 rays per scene (camera, sphere-to-box, interior with
    finite t_max, shadow, adversarial)."""
    scenes = [("shadows", T.scenes.shadows(resolution=256)[:2])]
    if os.path.exists(T.scenes.ASSET_PLY):
        scenes.append(("caustic-glass", T.scenes.caustic_glass(resolution=256)[:2]))
    scenes.append(("tess-60k", T.scenes.tessellated(cells=140, stacks=52, slices=50, res=(480, 270))[:2]))
    for label, (scene, camera) in scenes:
        ctx.upload(scene)
        total = 0
        for name, (o, d, tmax) in ray_sets(T, scene, camera, 2_400_000, 99).items():
            ctx.set_option("slab", 0)
            p0, t0, b0 = ctx.intersect(o, d, tmax)
            o0 = ctx.occluded(o, d, tmax)
            for variant in (2,):
                ctx.set_option("slab", variant)
                p2, t2, b2 = ctx.intersect(o, d, tmax)
                o2 = ctx.occluded(o, d, tmax)
                assert np.array_equal(p0, p2), f"{label}/{name}/slab{variant}: {np.count_nonzero(p0 != p2)} primitive ids differ"
                assert np.array_equal(t0.view(np.uint32), t2.view(np.uint32)) and np.array_equal(b0.view(np.uint32), b2.view(np.uint32))
                assert np.array_equal(o0, o2), f"{label}/{name}/slab{variant}: {np.count_nonzero(o0 != o2)} any-hit results differ"
            ctx.set_option("slab", 2)
            total += len(o)
        print(f"guarded == literal on {total} rays of {label}")
        assert total >= 10_000_000


# ---------------------------------------------------------------- image parity
def whitted_pair(T, ctx, scene, camera, spp, depth, seed=7):
    flat = ctx.upload(scene)
    osc = oracle_lib.OracleScene(flat)
    film = camera.film
    cam, fd = camera.pod(), film.desc()
    gpu = np.zeros_like(film.pixels)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), spp, depth, C.c_uint64(seed), T._lib.ptr(gpu)))
    st = ctx.stats()
    ref = np.zeros_like(film.pixels)
    cnt = osc.render_whitted(cam, fd, spp, depth, seed, ref)
    return gpu, ref, st, cnt


def image_report(gpu, ref, label):
    scale = float(np.abs(ref[..., :3]).max())
    err = np.abs(gpu - ref)
    rel_mse = float(np.mean((gpu[..., :3] - ref[..., :3]) ** 2) / max(1e-20, np.mean(ref[..., :3] ** 2)))
    frac_close = float(np.mean(np.all(err[..., :3] <= 2e-3 * scale + 1e-6, axis=-1)))
    werr = float(np.abs(gpu[..., 3] - ref[..., 3]).max() / max(1e-12, np.abs(ref[..., 3]).max()))
    print(f"image {label}: relMSE {rel_mse:.3e}  pixels within 0.2% of peak {frac_close:.5f}  weight max rel err {werr:.2e}")
    return rel_mse, frac_close, werr


def test_whitted_shadows_image(T, ctx):
    scene, camera, _ = T.scenes.shadows(resolution=96)
    gpu, ref, st, cnt = whitted_pair(T, ctx, scene, camera, 4, 5)
    rel_mse, frac, werr = image_report(gpu, ref, "whitted/shadows")
    # tolerance: same RNG, same arithmetic up to libm ULPs and float-add order -> near-identical films
    assert werr < 1e-5 and rel_mse < 1e-6 and frac > 0.999
    assert st["rays_extend"] == int(cnt[0]) and st["rays_shadow"] == int(cnt[1])


def test_whitted_tessellated_image(T, ctx):
    scene, camera, _ = T.scenes.tessellated(cells=48, stacks=26, slices=24, res=(160, 90))
    gpu, ref, st, cnt = whitted_pair(T, ctx, scene, camera, 4, 5)
    rel_mse, frac, werr = image_report(gpu, ref, "whitted/tess-small")
    assert werr < 1e-5 and rel_mse < 1e-5 and frac > 0.998
    assert abs(st["rays_extend"] - int(cnt[0])) <= 2e-4 * int(cnt[0])
    assert float(ref[..., 1].max()) > 0


def test_whitted_thin_lens_camera(T, ctx):
    """lens_radius > 0: the depth-of-field branch of generate_ray (perspective.jl:94-103) with lens samples from the
    shared counter-based RNG (dimensions 2-3)."""
    scene, _, _ = T.scenes.shadows(resolution=64)
    film = T.Film([64, 64], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    # the reference's camera looks down -z (look_at: z_axis = normalize(position - target)), so ray.d[3] < 0 and the
    # plane of focus t = focal_distance / ray.d[3] lies in front of the camera only for a NEGATIVE focal_distance
    camera = T.PerspectiveCamera(T.look_at([0, 15, 50], [0, 0, -2], [0, 1, 0]), T.Bounds2([-1, -1], [1, 1]), 0.0, 1.0,
                                 0.6, -54.0, 90.0, film)
    gpu, ref, st, cnt = whitted_pair(T, ctx, scene, camera, 8, 4)
    rel_mse, frac, werr = image_report(gpu, ref, "whitted/thin-lens")
    assert werr < 1e-5 and rel_mse < 1e-6 and frac > 0.999 and float(ref[..., 1].max()) > 0
    # and it really blurs: differs from the pinhole render
    pin = T.PerspectiveCamera(T.look_at([0, 15, 50], [0, 0, -2], [0, 1, 0]), T.Bounds2([-1, -1], [1, 1]), 0.0, 1.0, 0.0, 54.0, 90.0, film)
    gpu2, _, _, _ = whitted_pair(T, ctx, scene, pin, 8, 4)
    assert np.abs(gpu2[..., :3] - gpu[..., :3]).max() > 1e-2 * np.abs(gpu2[..., :3]).max()


def test_whitted_accumulates_into_film(T, ctx):
    """The reference never clears the film (merge_film_tile! adds, Q14): a second render doubles xyz and weights."""
    scene, camera, _ = T.scenes.shadows(resolution=48)
    ctx.upload(scene)
    cam, fd = camera.pod(), camera.film.desc()
    a = np.zeros_like(camera.film.pixels)
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 2, 3, C.c_uint64(3), T._lib.ptr(a)))
    b = a.copy()
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 2, 3, C.c_uint64(3), T._lib.ptr(b)))
    assert np.allclose(b, 2 * a, rtol=1e-5, atol=1e-7)


def test_whitted_queue_overflow_retry(T, ctx):
    """Glass spawns a reflected and a transmitted ray per hit (sampler.jl:95-98), so a bounce level can hold more rays
    than the batch has samples.  With the queue capacity squeezed to 100 % of the batch the queues overflow: the batch
    is not splatted and is re-run in halves - same image as the roomy run, and the oracle agrees."""
    glass = T.GlassMaterial(T.ConstantTexture(T.RGBSpectrum(1.0)), T.ConstantTexture(T.RGBSpectrum(1.0)), T.ConstantTexture(0.0),
                            T.ConstantTexture(0.0), T.ConstantTexture(1.5), True)
    white = T.MatteMaterial(T.ConstantTexture(T.RGBSpectrum(1.0)), T.ConstantTexture(0.0))
    prims = [T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0.5, 0.5, -2.5]), False), 0.55, 360.0), glass)]
    tris = T.create_triangle_mesh(T.ShapeCore(T.translate([0, 0, -2]), False), 2, [1, 2, 3, 1, 4, 3], 4,
                                  [[-2, -0.2, 2], [-2, -0.2, -3], [3, -0.2, -3], [3, -0.2, 2]], [[0, 1, 0]] * 4)
    prims += [T.GeometricPrimitive(t, white) for t in tris]
    scene = T.Scene([T.PointLight(T.translate([-1, 3, 0]), T.RGBSpectrum(25.0))], T.BVHAccel(prims, 1))
    film = T.Film([64, 64], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    camera = T.PerspectiveCamera(T.look_at([0, 15, 50], [0, 0, -2], [0, 1, 0]), T.Bounds2([-1, -1], [1, 1]), 0, 1, 0, 1e6, 90.0, film)
    flat = ctx.upload(scene)
    cam, fd = camera.pod(), camera.film.desc()
    a = np.zeros_like(camera.film.pixels)
    ctx.set_option("batch", 1 << 26)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 4, 8, C.c_uint64(5), T._lib.ptr(a)))
    n_over = ctx.stats()["queue_overflows"]
    b = np.zeros_like(a)
    ctx.set_option("cap_percent", 100)      # 25 600 slots per batch < 25 830 rays at bounce level 2
    ctx.set_option("batch", 8192)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 4, 8, C.c_uint64(5), T._lib.ptr(b)))
    n_over_small = ctx.stats()["queue_overflows"]
    ctx.set_option("batch", 1 << 26)
    ctx.set_option("cap_percent", 200)
    print("queue overflows: roomy", n_over, " squeezed", n_over_small)
    assert n_over == 0 and n_over_small > 0
    assert np.allclose(a, b, rtol=2e-4, atol=1e-6)
    ref = np.zeros_like(a)
    oracle_lib.OracleScene(flat).render_whitted(cam, fd, 4, 8, 5, ref)
    rel_mse, frac, werr = image_report(a, ref, "whitted/glass-ball depth 8")
    # tolerance: the analytic sphere's hit record goes through acosf / sinf (sphere.jl:92,152); last-ULP libm differences
    # are amplified by 8 specular bounces inside the ball (measured relMSE 6.9e-6, 99.8 % of pixels within 0.2 % of peak)
    assert rel_mse < 1e-4 and frac > 0.99 and werr < 1e-5


def sppm_pair(T, ctx, scene, camera, r0, depth, iters, photons, seed=11):
    flat = ctx.upload(scene)
    osc = oracle_lib.OracleScene(flat)
    film = camera.film
    cam, fd = camera.pod(), film.desc()
    h, w = film.pixels.shape[:2]
    gpu = np.zeros((h, w, 3), np.float32)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_sppm(ctx.h, C.byref(cam), C.byref(fd), r0, depth, iters, photons, 0, C.c_uint64(seed),
                                        C.cast(None, T._lib.SPPM_CB), None, T._lib.ptr(gpu)))
    st = ctx.stats()
    ref = np.zeros_like(gpu)
    cnt = osc.render_sppm(cam, fd, r0, depth, iters, photons, seed, ref)
    return gpu, ref, st, cnt


def test_sppm_shadows_image(T, ctx):
    scene, camera, kw = T.scenes.shadows(resolution=96)
    gpu, ref, st, cnt = sppm_pair(T, ctx, scene, camera, 0.025, 5, 4, -1)
    g4 = np.concatenate([gpu, np.ones_like(gpu[..., :1])], -1)
    r4 = np.concatenate([ref, np.ones_like(ref[..., :1])], -1)
    rel_mse, frac, _ = image_report(g4, r4, "sppm/shadows")
    print("sppm rays gpu", st["rays_extend"], st["rays_shadow"], "oracle", cnt, "deposits", st["sppm_deposits"])
    # tolerance (per scene, north_star): flux atomics reorder float adds and libm ULPs can flip a rare Russian-roulette
    # or Fresnel choice, moving a single photon path
    assert rel_mse < 5e-3 and frac > 0.98
    assert abs(st["rays_extend"] - int(cnt[0])) <= 2e-3 * int(cnt[0])


def test_sppm_caustic_glass_image(T, ctx):
    if not os.path.exists(T.scenes.ASSET_PLY):
        pytest.skip("asset missing")
    scene, camera, kw = T.scenes.caustic_glass(resolution=64)
    gpu, ref, st, cnt = sppm_pair(T, ctx, scene, camera, 0.075, 5, 3, 20000)
    g4 = np.concatenate([gpu, np.ones_like(gpu[..., :1])], -1)
    r4 = np.concatenate([ref, np.ones_like(ref[..., :1])], -1)
    rel_mse, frac, _ = image_report(g4, r4, "sppm/caustic-glass")
    print("sppm rays gpu", st["rays_extend"], st["rays_shadow"], "oracle", cnt, "deposits", st["sppm_deposits"])
    # Per-scene tolerance (north_star): the spot light's photon directions go through sinf/cosf, whose last-ULP
    # differences (CUDA libm vs glibc) move a handful of photons across a triangle edge of the 88k-triangle glass mesh;
    # refraction then sends them elsewhere, and in the caustic (dense visible points) one photon touches ~30 pixels.
    # Measured: 21 of 171 824 rays differ, relMSE 3.2e-3, 92 % of pixels within 0.2 % of peak.
    assert rel_mse < 2e-2 and frac > 0.85
    assert abs(st["rays_extend"] - int(cnt[0])) <= 1e-3 * int(cnt[0])
    assert float(ref.max()) > 0


def _appearance_scene(T, res=72):
    """Rough glass (microfacet reflection + transmission), Oren-Nayar matte, plastic, a DirectionalLight (preprocessed)
    next to a point light: the appearance surface SURVEY.md §8f.3 lists."""
    rough_glass = T.GlassMaterial(T.ConstantTexture(T.RGBSpectrum(1.0)), T.ConstantTexture(T.RGBSpectrum(0.9)),
                                  T.ConstantTexture(0.2), T.ConstantTexture(0.35), T.ConstantTexture(1.5), True)
    oren = T.MatteMaterial(T.ConstantTexture(T.RGBSpectrum(0.8, 0.6, 0.4)), T.ConstantTexture(35.0))
    plastic = T.PlasticMaterial(T.ConstantTexture(T.RGBSpectrum(0.3, 0.5, 0.3)), T.ConstantTexture(T.RGBSpectrum(0.4)),
                                T.ConstantTexture(0.15), True)
    prims = [T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0.3, 0.25, -2.4]), False), 0.25, 360.0), rough_glass),
             T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0.75, 0.2, -2.5]), False), 0.2, 360.0), plastic)]
    tris = T.create_triangle_mesh(T.ShapeCore(T.translate([0, 0, -2]), False), 4, [1, 2, 3, 1, 4, 3, 2, 3, 5, 6, 5, 3], 6,
                                  [[0, 0, 0], [0, 0, -1], [1, 0, -1], [1, 0, 0], [0, 1, -1], [1, 1, -1]],
                                  [[0, 1, 0]] * 4 + [[0, 0, 1]] * 2)
    prims += [T.GeometricPrimitive(t, oren) for t in tris]
    scene = T.Scene([T.PointLight(T.translate([-1, 1, 0]), T.RGBSpectrum(10.0))], T.BVHAccel(prims, 1))
    sun = T.DirectionalLight(T.rotate_x(20.0), T.RGBSpectrum(1.5, 1.4, 1.2), [0.3, 1.0, 0.6])
    sun.preprocess(scene)
    scene.lights.append(sun)
    film = T.Film([res, res], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    camera = T.PerspectiveCamera(T.look_at([0, 15, 50], [0, 0, -2], [0, 1, 0]), T.Bounds2([-1, -1], [1, 1]), 0, 1, 0, 1e6, 90.0, film)
    return scene, camera


def test_whitted_appearance_surface(T, ctx):
    scene, camera = _appearance_scene(T)
    gpu, ref, st, cnt = whitted_pair(T, ctx, scene, camera, 4, 5)
    rel_mse, frac, werr = image_report(gpu, ref, "whitted/rough-glass+oren-nayar+directional")
    assert float(ref[..., 1].max()) > 0
    assert werr < 1e-5 and rel_mse < 1e-5 and frac > 0.995
    assert st["rays_extend"] == int(cnt[0]) and st["rays_shadow"] == int(cnt[1])


def test_sppm_appearance_surface(T, ctx):
    """Same materials through the SPPM passes (sample_f of the microfacet lobes, Oren-Nayar visible points); the
    directional light is dropped: the reference has no sample_le for it and both sides refuse it."""
    scene, camera = _appearance_scene(T, 64)
    with_sun = scene
    scene = T.Scene(with_sun.lights[:1], with_sun.aggregate)
    gpu, ref, st, cnt = sppm_pair(T, ctx, scene, camera, 0.04, 5, 3, 60000)
    g4 = np.concatenate([gpu, np.ones_like(gpu[..., :1])], -1)
    r4 = np.concatenate([ref, np.ones_like(ref[..., :1])], -1)
    rel_mse, frac, _ = image_report(g4, r4, "sppm/rough-glass+oren-nayar")
    print("sppm rays gpu", st["rays_extend"], st["rays_shadow"], "oracle", cnt, "deposits", st["sppm_deposits"])
    assert rel_mse < 2e-2 and frac > 0.9 and float(ref.max()) > 0
    ctx.upload(with_sun)
    cam, fd = camera.pod(), camera.film.desc()
    rc = ctx.lib.trace_sppm_begin(ctx.h, C.byref(cam), C.byref(fd), 0.04, 5, 1000, C.c_uint64(1))
    assert rc != 0 and b"DirectionalLight" in ctx.lib.trace_last_error(ctx.h)


def test_integrator_functors(T, ctx, tmp_path):
    """The reference-facing call: integrator(scene) renders into camera.film and saves a PNG."""
    scene, camera, kw = T.scenes.shadows(resolution=40, filename=str(tmp_path / "w.png"))
    img = T.WhittedIntegrator(camera, T.UniformSampler(2), 4, context=ctx)(scene)
    assert img.shape == (40, 40, 3) and float(img.max()) > 0 and os.path.getsize(tmp_path / "w.png") > 100
    scene, camera, kw = T.scenes.shadows(resolution=40, filename=str(tmp_path / "s.png"))
    img = T.SPPMIntegrator(camera, 0.05, 5, 2, -1, 2, context=ctx)(scene)
    assert img.shape == (40, 40, 3) and float(img.max()) > 0 and os.path.getsize(tmp_path / "s.png") > 100


def test_optin_sah_tree_on_gpu(T, ctx):
    """the opt-in conventional SAH tree (BVHAccel(..., builder="sah")).  On the GPU, over 2e6 rays of the
    caustic mesh: t bit-identical to the reference tree's, the same primitive except on ties of equal t, fewer box
    tests; the oracle walking the SAME sah tree agrees with the GPU bit for bit (prim, t, barycentrics); and a Whitted
    image rendered on either tree is the same image."""
    rng = np.random.default_rng(11)
    n = 2_000_000
    flats = {}
    for builder in ("reference", "sah"):
        scene, _, _ = T.scenes.caustic_glass(builder=builder)
        flats[builder] = (scene, scene.flatten())
    lo, hi = flats["reference"][1].nodes[0]["bmin"], flats["reference"][1].nodes[0]["bmax"]
    target = (lo + rng.random((n, 3), dtype=np.float32) * (hi - lo)).astype(np.float32)
    origin = np.where(rng.random((n, 1)) < 0.5, np.array([[0, 150, 150]], np.float32),
                      (target + rng.normal(size=(n, 3)).astype(np.float32) * 20)).astype(np.float32)
    d = (target - origin).astype(np.float32)
    out = {}
    for builder, (scene, flat) in flats.items():
        ctx.upload(scene)
        ctx.set_option("count_nodes", 1)
        ctx.reset_stats()
        out[builder] = ctx.intersect(origin, d) + (ctx.stats()["nodes_visited"],)
        ctx.set_option("count_nodes", 0)
    (pa, ta, ba, na), (pb, tb, bb, nb) = out["reference"], out["sah"]
    assert (pa != 0).sum() > n // 4
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
    assert (pa != pb).mean() < 1e-3
    assert nb < na
    m = 100_000
    op, ot, ob = oracle_lib.OracleScene(flats["sah"][1]).intersect(origin[:m], d[:m], slab=0)
    assert np.array_equal(op, pb[:m]) and np.array_equal(ot.view(np.uint32), tb[:m].view(np.uint32))
    assert np.array_equal(ob.view(np.uint32), bb[:m].view(np.uint32))
    # images
    films = []
    for builder in ("reference", "sah"):
        scene, camera, _ = T.scenes.tessellated(cells=48, stacks=26, slices=24, res=(160, 90), builder=builder)
        ctx.upload(scene)
        cam, fd = camera.pod(), camera.film.desc()
        film = np.zeros_like(camera.film.pixels)
        ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 4, 5, C.c_uint64(9), T._lib.ptr(film)))
        films.append(film)
    rel_mse, frac, werr = image_report(films[1], films[0], "whitted/tess-small sah-vs-reference tree")
    assert werr < 1e-5 and rel_mse < 1e-5 and frac > 0.998

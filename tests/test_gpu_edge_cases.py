"""Edge cases of the C ABI on the GPU (-m gpu): cropped / non-square films, several lights, empty and degenerate
scenes, bad arguments (nonzero return + message, never a crash), SPPM callback cadence."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu


def _whitted_both(T, ctx, scene, camera, spp=2, depth=4, seed=5):
    flat = ctx.upload(scene)
    cam, fd = camera.pod(), camera.film.desc()
    g = np.zeros_like(camera.film.pixels)
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), spp, depth, C.c_uint64(seed), T._lib.ptr(g)))
    r = np.zeros_like(g)
    oracle_lib.OracleScene(flat).render_whitted(cam, fd, spp, depth, seed, r)
    return g, r


def test_cropped_non_square_film_two_lights(T, ctx):
    """Film crop window (film.jl:41-44) that is neither full nor square, filter radius 2, two lights."""
    scene, _, _ = T.scenes.shadows(resolution=64)
    scene = T.Scene(scene.lights + [T.SpotLight(T.translate([0.5, 2.0, -2.0]) * T.rotate_x(90.0), T.RGBSpectrum(8.0, 6.0, 4.0), 40.0, 25.0)],
                    scene.aggregate)
    film = T.Film([96, 64], T.Bounds2([0.25, 0.1], [0.8, 0.95]), T.LanczosSincFilter([2, 2], 3.0), 1.0, 1.0, None)
    camera = T.PerspectiveCamera(T.look_at([0, 15, 50], [0, 0, -2], [0, 1, 0]), T.Bounds2([-1.5, -1], [1.5, 1]), 0, 1, 0, 1e6, 90.0, film)
    assert film.pixels.shape[:2] != (64, 96) and film.pixels.shape[0] != film.pixels.shape[1]
    g, r = _whitted_both(T, ctx, scene, camera)
    assert float(r[..., 3].max()) > 0 and float(r[..., 1].max()) > 0
    assert np.allclose(g[..., 3], r[..., 3], rtol=1e-5, atol=1e-7)
    scale = np.abs(r[..., :3]).max()
    assert np.abs(g[..., :3] - r[..., :3]).max() <= 2e-4 * scale


def test_scene_that_misses_everything_and_no_lights(T, ctx):
    """Every ray misses (camera looks away): zero radiance, but the filter weights are still accumulated; a scene
    without lights renders black and SPPM refuses it."""
    scene, _, _ = T.scenes.shadows(resolution=32)
    film = T.Film([32, 32], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    away = T.PerspectiveCamera(T.look_at([0, 15, 50], [0, 30, 102], [0, 1, 0]), T.Bounds2([-1, -1], [1, 1]), 0, 1, 0, 1e6, 90.0, film)
    g, r = _whitted_both(T, ctx, scene, away)
    assert np.all(g[..., :3] == 0) and np.allclose(g[..., 3], r[..., 3], rtol=1e-5) and g[..., 3].min() > 0
    dark = T.Scene([], scene.aggregate)
    cam0 = T.scenes.shadows(resolution=32)[1]
    g, r = _whitted_both(T, ctx, dark, cam0)
    assert np.all(g[..., :3] == 0) and np.all(r[..., :3] == 0)
    cam, fd = cam0.pod(), cam0.film.desc()
    rgb = np.zeros((32, 32, 3), np.float32)
    rc = ctx.lib.trace_render_sppm(ctx.h, C.byref(cam), C.byref(fd), 0.05, 3, 1, -1, 0, C.c_uint64(1), C.cast(None, T._lib.SPPM_CB), None, T._lib.ptr(rgb))
    assert rc != 0 and b"no lights" in ctx.lib.trace_last_error(ctx.h)


def test_single_primitive_and_all_degenerate_scenes(T, ctx):
    mat = T.MatteMaterial(T.ConstantTexture(T.RGBSpectrum(0.7)), T.ConstantTexture(0.0))
    one = T.create_triangle_mesh(T.ShapeCore(T.Transformation(), False), 1, [1, 2, 3], 3, [[0, 0, -3], [1, 0, -3], [0.5, 1, -3.2]])
    scene = T.Scene([T.PointLight(T.translate([0.5, 0.5, 0]), T.RGBSpectrum(5.0))], T.BVHAccel([T.GeometricPrimitive(one[0], mat)]))
    flat = ctx.upload(scene)
    o = np.array([[0.5, 0.3, 0.0], [5, 5, 0]], np.float32)
    d = np.array([[0.01, 0.02, -1.0], [0, 0, -1]], np.float32)
    prim, t, b = ctx.intersect(o, d)
    rp, rt, rb = oracle_lib.OracleScene(flat).intersect(o, d)
    assert np.array_equal(prim, rp) and np.array_equal(t.view(np.uint32), rt.view(np.uint32)) and prim[0] == 1 and prim[1] == 0
    # only degenerate triangles (is_degenerate, triangle_mesh.jl:65-68): nothing is ever hit
    deg = T.create_triangle_mesh(T.ShapeCore(T.Transformation(), False), 2, [1, 2, 3, 1, 1, 2], 3, [[0, 0, -3], [1, 0, -3], [2, 0, -3]])
    scene = T.Scene([], T.BVHAccel([T.GeometricPrimitive(x, mat) for x in deg]))
    ctx.upload(scene)
    prim, _, _ = ctx.intersect([[0.5, 1, 0], [0.5, 0, 0]], [[0, -0.3, -1], [0, 0, -1]])
    assert np.all(prim == 0) and not ctx.occluded([[0.5, 0, 0]], [[0, 0, -1]])[0]


def test_bad_arguments_fail_cleanly(T, ctx):
    lib = ctx.lib
    scene, camera, _ = T.scenes.shadows(resolution=32)
    ctx.upload(scene)
    cam, fd = camera.pod(), camera.film.desc()
    film = np.zeros_like(camera.film.pixels)
    assert lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 0, 5, C.c_uint64(1), T._lib.ptr(film)) != 0     # spp = 0
    assert lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 1, 0, C.c_uint64(1), T._lib.ptr(film)) != 0     # depth = 0
    assert lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 1, 5, C.c_uint64(1), None) != 0                 # null film
    assert lib.trace_set_option(ctx.h, b"no_such_option", 1) != 0 and b"unknown option" in lib.trace_last_error(ctx.h)
    assert lib.trace_set_option(ctx.h, b"slab", 7) != 0
    assert lib.trace_sppm_photon_pass(ctx.h, 1, 0, 10) != 0                                                           # no trace_sppm_begin
    bad = T._lib.FilmDesc()
    bad.crop_x0, bad.crop_y0, bad.crop_x1, bad.crop_y1 = 5, 5, 1, 1
    bad.filter_radius[0] = bad.filter_radius[1] = 1.0
    assert lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(bad), 1, 2, C.c_uint64(1), T._lib.ptr(film)) != 0    # empty crop
    # a corrupt scene description is rejected at upload, the previous scene stays usable
    flat = scene.flatten()
    d = flat.desc()
    nodes = flat.nodes.copy()
    nodes[0]["offset"] = 10_000
    d.nodes = T._lib.ptr(nodes)
    assert lib.trace_scene_upload(ctx.h, C.byref(d)) != 0 and b"out of range" in lib.trace_last_error(ctx.h)
    ctx._scene = None
    ctx.upload(scene)
    prim, _, _ = ctx.intersect([[0.5, 0.5, 5]], [[0.001, 0.001, -1]])
    assert prim.shape == (1,)


def test_sppm_callback_cadence_and_default_photons(T, ctx):
    """on_image fires every write_frequency iterations and after the last one (sppm.jl:167-171); photons default to
    area(crop_bounds) = (W-1)(H-1) (Q22)."""
    scene, camera, _ = T.scenes.shadows(resolution=24)
    ctx.upload(scene)
    cam, fd = camera.pod(), camera.film.desc()
    seen = []
    cb = T._lib.SPPM_CB(lambda user, it, ptr: seen.append(int(it)))
    rgb = np.zeros((24, 24, 3), np.float32)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_sppm(ctx.h, C.byref(cam), C.byref(fd), 0.05, 3, 7, -1, 3, C.c_uint64(1), cb, None, T._lib.ptr(rgb)))
    assert seen == [3, 6, 7]
    st = ctx.stats()
    # 7 iterations x (24^2 camera paths + 23^2 photons) rays at depth 1, plus bounces
    assert st["rays_extend"] >= 7 * (24 * 24 + 23 * 23)
    assert np.isfinite(rgb).all() and rgb.max() > 0


def test_sppm_pipeline_depth_and_call_granularity_do_not_change_the_image(T):
    """Iterations in flight (option sppm_pipeline: own visible-point arrays, queues and streams per slot; the camera pass no
    longer reads the radius) and the way the iterations are handed over (one trace_sppm_iterate call with look-ahead, one
    call per iteration, or the stepwise camera / grid / photon / update entry points) are scheduling only: after the same
    iterations every variant must show the image of the strictly serial run, up to the order of the float flux atomics -
    and that one is held to the oracle."""
    from trace_jl_b200 import distributed as D
    scene, camera, kw = T.scenes.shadows(resolution=96)
    iters, r0, depth = 7, kw["initial_search_radius"], kw["max_depth"]

    def run(pipeline, mode):
        ctx = T.Context(0)
        ctx.set_option("sppm_pipeline", pipeline)
        flat = ctx.upload(scene)
        sess = D.SPPMSession(ctx, scene, camera, r0, depth)
        if mode == "one call":
            sess.step(iters)
        elif mode == "per iteration":
            for _ in range(iters):
                sess.step(1)
        else:                                                  # stepwise entry points (slot 0 only)
            photons = int(camera.film.crop_bounds.area())
            for it in range(1, iters + 1):
                ctx.check(ctx.lib.trace_sppm_camera_pass(ctx.h, it))
                ctx.check(ctx.lib.trace_sppm_photon_pass(ctx.h, it, 0, photons))
                ctx.check(ctx.lib.trace_sppm_update(ctx.h))
            sess.iteration = iters
        img = sess.image()
        sess.close()
        ctx.close()
        return img, flat

    base, flat = run(1, "one call")
    assert np.isfinite(base).all() and base.max() > 0
    for pipeline, mode in ((4, "one call"), (8, "one call"), (2, "per iteration"), (4, "per iteration"), (4, "stepwise")):
        img, _ = run(pipeline, mode)
        assert np.allclose(img, base, rtol=2e-4, atol=1e-6), (pipeline, mode, float(np.abs(img - base).max()))
    cam, fd = camera.pod(), camera.film.desc()
    ref = np.zeros_like(base)
    oracle_lib.OracleScene(flat).render_sppm(cam, fd, r0, depth, iters, -1, 0x5EED0001, ref)
    rel = float(np.mean((base - ref) ** 2) / max(1e-12, np.mean(ref ** 2)))
    assert rel < 5e-3, rel


def test_whitted_graph_replay_follows_seed_and_camera(T):
    """The render is captured once into a CUDA graph and replayed; seed and camera are read through a device block, so
    a replay with another seed / camera must give what direct launches give (bit for bit: same kernels, same order of
    the film atomics only within float rounding)."""
    def render(ctx, seed, fov):
        scene, cam0, _ = T.scenes.shadows(resolution=96)
        film = T.scenes._film(96)
        window = T.Bounds2(T.Point2f(-1.0, -1.0), T.Point2f(1.0, 1.0))
        camera = T.PerspectiveCamera(cam0.camera_to_world, window, 0.0, 1.0, 0.0, 1e6, fov, film)
        ctx.upload(scene)
        cam, fd = camera.pod(), film.desc()
        ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 4, 5, C.c_uint64(seed), T._lib.ptr(film.pixels)))
        return film.pixels.copy()

    direct, graph = T.Context(0), T.Context(0)
    direct.set_option("graph", 0)
    graph.set_option("graph", 1)
    try:
        for seed, fov in ((1, 45.0), (2, 45.0), (2, 60.0), (1, 45.0)):
            a, b = render(direct, seed, fov), render(graph, seed, fov)
            assert a[..., 3].max() > 0
            assert np.allclose(a, b, rtol=1e-5, atol=1e-6), (seed, fov, float(np.abs(a - b).max()))
        assert not np.allclose(render(graph, 1, 45.0), render(graph, 2, 45.0), rtol=1e-5, atol=1e-6)
    finally:
        direct.close()
        graph.close()

"""Full-size GPU checks (-m gpu) at BASELINE.json's sizes, through size-independent properties (the oracle would take
minutes there): tess-1M (999 840 triangles, 1920x1080) and the shadows SPPM scene at 1024x1024."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tess(T):
    scene, camera, kw = T.scenes.tessellated()          # C3: 600^2 heightfield cells + two 266x264 UV spheres
    assert scene.aggregate.n_primitives == 999_840
    return scene, camera


def test_tess1m_query_properties(T, ctx, tess):
    scene, camera = tess
    flat = ctx.upload(scene)
    rng = np.random.default_rng(1)
    n = 2_000_000
    lo, hi = np.array(flat.nodes[0]["bmin"]), np.array(flat.nodes[0]["bmax"])
    o = rng.uniform(lo - 5, hi + 5, (n, 3)).astype(np.float32)
    d = (rng.uniform(lo, hi, (n, 3)) - o).astype(np.float32)
    ctx.set_option("slab", 0)
    p0, t0, b0 = ctx.intersect(o, d)
    occ0 = ctx.occluded(o, d)
    ctx.set_option("slab", 2)
    p2, t2, b2 = ctx.intersect(o, d)
    occ2 = ctx.occluded(o, d)
    # guarded == literal, bit for bit
    assert np.array_equal(p0, p2) and np.array_equal(t0.view(np.uint32), t2.view(np.uint32)) and np.array_equal(b0.view(np.uint32), b2.view(np.uint32))
    assert np.array_equal(occ0, occ2)
    # any-hit (t_max = Inf) <=> closest-hit found something
    assert np.array_equal(occ2, p2 != 0)
    hit = p2 != 0
    assert 0.2 < hit.mean() < 1.0
    # idempotence: the closest hit does not depend on t_max as long as t_max lies beyond it (the acceptance test
    # compares in scaled space, triangle_mesh.jl:211-214, so the margin is relative, not one ULP)
    tm = (t2[hit] * np.float32(1.001)).astype(np.float32)
    p3, t3, _ = ctx.intersect(o[hit], d[hit], tm)
    assert np.array_equal(p3, p2[hit]) and np.array_equal(t3.view(np.uint32), t2[hit].view(np.uint32))
    # ... and with t_max short of it that hit is gone: nothing or something farther is never reported closer
    tm = (t2[hit] * np.float32(0.999)).astype(np.float32)
    p4, t4, _ = ctx.intersect(o[hit], d[hit], tm)
    assert np.all(p4 == 0)
    # barycentrics of triangle hits are a partition of unity (watertight test: all edge functions share a sign)
    bb = b2[hit]
    assert np.all(bb >= -1e-6) and np.all(bb.sum(1) <= 1 + 1e-5)
    # original indices are valid and every reported t is positive
    assert p2.max() <= 999_840 and np.all(t2[hit] > 0)


def test_tess1m_whitted_linearity_and_weights(T, ctx, tess):
    """Radiance is linear in the light intensity (x2 is exact in binary floating point up to atomic-add order), and the
    filter-weight plane of the film depends on the camera samples only, not on the scene."""
    scene, camera = tess
    ctx.upload(scene)
    cam, fd = camera.pod(), camera.film.desc()
    a = np.zeros_like(camera.film.pixels)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 2, 5, C.c_uint64(3), T._lib.ptr(a)))
    st = ctx.stats()
    assert st["rays_extend"] >= 1922 * 1082 * 2 and st["rays_shadow"] > 0 and st["queue_overflows"] == 0
    assert np.isfinite(a).all() and float(a[..., 1].max()) > 0
    # brighter light
    scene2 = T.Scene([T.PointLight(scene.lights[0].light_to_world, T.RGBSpectrum(800.0))], scene.aggregate)
    scene2._flat = None
    ctx.upload(scene2)
    b = np.zeros_like(a)
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 2, 5, C.c_uint64(3), T._lib.ptr(b)))
    assert np.allclose(b[..., :3], 2 * a[..., :3], rtol=1e-4, atol=1e-7)
    assert np.allclose(b[..., 3], a[..., 3], rtol=1e-6)
    # weights: every interior pixel received samples; each sample splats into at most (2r + 2)^2 = 16 pixels (Q4)
    w = a[..., 3]
    assert w[8:-8, 8:-8].min() > 0
    # the film is accumulated into, never cleared (Q14)
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), 2, 5, C.c_uint64(3), T._lib.ptr(b)))
    assert np.allclose(b[..., :3], 4 * a[..., :3], rtol=1e-4, atol=1e-7) and np.allclose(b[..., 3], 2 * w, rtol=1e-6)
    ctx.upload(scene)


def test_shadows_1024_sppm_properties(T, ctx):
    """docs/code/spheres.jl at its shipped 1024x1024: radii only shrink, the image is finite and non-negative, and
    photon sharding (2 halves into one flux buffer) equals the unsharded pass."""
    from trace_jl_b200 import distributed as D
    import torch
    scene, camera, kw = T.scenes.shadows(resolution=1024)
    sess = D.SPPMSession(ctx, scene, camera, kw["initial_search_radius"], kw["max_depth"])
    assert sess.photons == 1023 * 1023
    sess.step()
    img1 = sess.image()
    sess.step()
    img2 = sess.image()
    sess.close()
    assert np.isfinite(img1).all() and np.isfinite(img2).all() and img1.min() >= 0 and img2.min() >= 0
    assert float(img2.mean()) > 0.05
    # sharded photon pass: iteration 1 traced as two halves must deposit the same flux (up to float-add order)
    sess = D.SPPMSession(ctx, scene, camera, kw["initial_search_radius"], kw["max_depth"])
    ctx.check(ctx.lib.trace_sppm_camera_pass(ctx.h, 1))
    half = sess.photons // 2
    ctx.check(ctx.lib.trace_sppm_photon_pass(ctx.h, 1, 0, half))
    ctx.check(ctx.lib.trace_sppm_photon_pass(ctx.h, 1, half, sess.photons))
    ctx.check(ctx.lib.trace_sppm_update(ctx.h))
    sess.iteration = 1
    img_sharded = sess.image()
    sess.close()
    assert np.allclose(img_sharded, img1, rtol=2e-4, atol=1e-6)


def test_sppm_row_sharding_protocol_on_one_gpu(T):
    """The multi-GPU SPPM protocol (rows of the camera pass dealt round-robin, all-gather of the visible points,
    photon ranges, all-reduce of the flux, all-gather of Ld) emulated with TWO contexts on one GPU: the exchanged
    slices are copied between the contexts' buffers by hand.  The image must equal the unsharded render."""
    import torch
    from trace_jl_b200 import distributed as D
    scene, camera, kw = T.scenes.shadows(resolution=151)        # 151 rows over 3 ranks: padding rows in play
    iters, r0, depth = 3, 0.03, 5
    ref_ctx = T.Context(0)
    sess = D.SPPMSession(ref_ctx, scene, camera, r0, depth)
    for _ in range(iters):
        sess.step()
    want = sess.image()
    sess.close()
    ref_ctx.close()

    world = 3
    ctxs = [T.Context(0) for _ in range(world)]
    # no process group: we play the collectives ourselves
    sessions = []
    for r, c in enumerate(ctxs):
        c.set_option("world", world)
        c.set_option("rank", r)
        cam, fd = camera.pod(), camera.film.desc()
        c.upload(scene)
        c.check(c.lib.trace_sppm_begin(c.h, C.byref(cam), C.byref(fd), r0, depth, -1, C.c_uint64(0x5EED0001)))
        bufs = []
        for which in range(7):
            n = C.c_int64()
            ptr = c.lib.trace_sppm_buffer_device(c.h, which, C.byref(n))
            bufs.append(torch.as_tensor(D._DevicePtr(ptr, n.value), device="cuda:0"))
        sessions.append(bufs)
    photons = int(camera.film.crop_bounds.area())

    def gather(which):
        n = sessions[0][which].numel() // world
        for dst in range(world):
            for src in range(world):
                if src != dst:
                    for c in ctxs:
                        c.synchronize()
                    sessions[dst][which][src * n:(src + 1) * n].copy_(sessions[src][which][src * n:(src + 1) * n])
        torch.cuda.synchronize()

    for it in range(1, iters + 1):
        for c in ctxs:
            c.check(c.lib.trace_sppm_camera_pass(c.h, it))
        for which in range(2, 7):
            gather(which)
        for c in ctxs:
            c.check(c.lib.trace_sppm_build_grid(c.h))
        for r, c in enumerate(ctxs):
            b, e = D.photon_range(photons, r, world)
            c.check(c.lib.trace_sppm_photon_pass(c.h, it, b, e))
        for c in ctxs:
            c.synchronize()
        total = sum(bufs[0] for bufs in sessions)
        torch.cuda.synchronize()
        for bufs in sessions:
            bufs[0].copy_(total)
        torch.cuda.synchronize()
        for c in ctxs:
            c.check(c.lib.trace_sppm_update(c.h))
    gather(1)
    got = np.zeros_like(want)
    ctxs[1].check(ctxs[1].lib.trace_sppm_image(ctxs[1].h, iters, T._lib.ptr(got)))
    for c in ctxs:
        c.check(c.lib.trace_sppm_end(c.h))
        c.close()
    assert np.isfinite(got).all() and float(want.max()) > 0
    assert np.allclose(got, want, rtol=3e-4, atol=1e-6), float(np.abs(got - want).max())

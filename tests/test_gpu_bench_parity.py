"""GPU-vs-oracle parity ON THE BENCHMARKED CONFIGURATIONS (-m gpu): the CUDA path, through the C ABI, against the CPU
oracle on the scenes, cameras, sample counts and depths BASELINE.json names - not on toy stand-ins.

  C3  tess-1M (999 840 triangles): 2.5e6 rays of five families bit-equal (every walk / slab variant), and the Whitted
      film of the real 1920x1080 x 16 spp x depth 5 camera over one rank-of-8's share of the 16x16 tiles
  C5  tess-10M (10 002 224 triangles): 2.1e5 rays bit-equal; the 4096^2 x 64 spp x depth 8 film over a 1/512 tile share
  C4  caustic_moving: TWO lights -> sample_discrete / uniform_sample_one_light with n > 1 (src/sampler/sampling.jl:3-41,
      src/integrators/sppm.jl:503-517), 3 iterations
  C1  caustic_glass at depth 5 (the shipped script) and 8 (README / BASELINE.json)
  C2 / C1 converged: the GPU image after n iterations against the oracle's image after 4n (per-scene tolerance, stated
      here and in BASELINE.md)
The oracle intersects ~250 k rays/s per thread, so these sizes cost seconds on the box's host cores.
"""
import ctypes as C
import os
import time

import numpy as np
import pytest

import oracle_lib
from test_gpu_parity import check_scene, image_report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tess1m(T):
    scene, camera, kw = T.scenes.tessellated()          # C3: 600^2 heightfield cells + two 266x264 UV spheres
    assert scene.aggregate.n_primitives == 999_840
    return scene, camera


def test_c3_tess1m_rays_bit_equal(T, ctx, tess1m):
    """2.5e6 rays (camera, sphere-to-box, interior with finite t_max, shadow, adversarial) on the 1M-triangle scene:
    primitive ids, t, barycentrics and any-hit booleans equal the oracle's bit for bit, for the literal and the guarded
    box test, the reference loop, the batched-leaf loop and the pair-node walk."""
    scene, camera = tess1m
    t0 = time.time()
    summary = check_scene(T, ctx, scene, camera, 500_000, 31, "tess-1M")
    n_rays = sum(s[3] for s in summary if s[2] == 0)
    print(f"tess-1M: {n_rays} rays per variant, {time.time() - t0:.1f} s")
    assert n_rays >= 2_500_000
    # the families are not vacuous: camera rays hit ~30 % (the rest is sky), interior / shadow rays mostly hit
    hits = {s[1]: s[4] / s[3] for s in summary if s[2] == 0}
    assert 0.2 < hits["R1_camera"] < 0.5 and hits["R3_interior"] > 0.3 and hits["R5_adversarial"] > 0.5


def _tile_share_films(T, ctx, scene, camera, spp, depth, seed, rank, world):
    """Film of the tiles k % world == rank: GPU (options rank / world select them) and oracle (explicit tile list)."""
    from trace_jl_b200 import distributed as D
    flat = ctx.upload(scene)
    osc = oracle_lib.OracleScene(flat)
    cam, fd = camera.pod(), camera.film.desc()
    ctx.set_option("world", world)
    ctx.set_option("rank", rank)
    gpu = np.zeros_like(camera.film.pixels)
    ctx.reset_stats()
    try:
        ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), spp, depth, C.c_uint64(seed), T._lib.ptr(gpu)))
    finally:
        ctx.set_option("rank", 0)
        ctx.set_option("world", 1)
    st = ctx.stats()
    ref = np.zeros_like(gpu)
    tiles = D.tile_shard(D.n_sample_tiles(camera.film), rank, world)
    cnt = osc.render_whitted_tiles(cam, fd, spp, depth, seed, ref, tiles)
    return gpu, ref, st, cnt, len(tiles)


def test_c3_whitted_film_on_the_real_camera(T, ctx, tess1m):
    """The headline configuration itself - 1920x1080, 16 spp, depth 5 - over rank 3 of 8's tiles (1 029 of 8 228, spread
    over the whole image): the film (X, Y, Z, weight) against the oracle's, and the ray counts."""
    scene, camera = tess1m
    gpu, ref, st, cnt, n_tiles = _tile_share_films(T, ctx, scene, camera, 16, 5, 77, 3, 8)
    rel_mse, frac, werr = image_report(gpu, ref, f"whitted/tess-1M 1920x1080x16spp depth 5, {n_tiles} tiles")
    assert float(ref[..., 1].max()) > 0 and np.count_nonzero(ref[..., 3]) > 200_000
    # tolerance: same per-sample RNG, same hits; what differs is libm ULPs in shading and the order of the film's float adds
    assert werr < 1e-5 and rel_mse < 1e-6 and frac > 0.999
    assert abs(st["rays_extend"] - int(cnt[0])) <= 1e-4 * int(cnt[0]) and abs(st["rays_shadow"] - int(cnt[1])) <= 1e-4 * int(cnt[1])
    assert st["queue_overflows"] == 0


@pytest.fixture(scope="module")
def tess10m(T):
    t0 = time.time()
    scene, camera, kw = T.scenes.tessellated(cells=1900, stacks=835, slices=834, res=(4096, 4096), window=((-50.0, -50.0), (50.0, 50.0)))
    assert scene.aggregate.n_primitives == 10_002_224
    scene.flatten()
    print(f"tess-10M built in {time.time() - t0:.1f} s")
    return scene, camera


def test_c5_tess10m_rays_bit_equal(T, ctx, tess10m):
    scene, camera = tess10m
    summary = check_scene(T, ctx, scene, camera, 42_000, 32, "tess-10M")
    assert sum(s[3] for s in summary if s[2] == 0) >= 200_000


def test_c5_whitted_film_on_the_real_camera(T, ctx, tess10m):
    """4096x4096, 64 spp, depth 8 on the 10M-triangle scene, over a 1/512 share of the tiles (129 tiles, 2.1e6 samples)."""
    scene, camera = tess10m
    gpu, ref, st, cnt, n_tiles = _tile_share_films(T, ctx, scene, camera, 64, 8, 78, 5, 512)
    rel_mse, frac, werr = image_report(gpu, ref, f"whitted/tess-10M 4096^2x64spp depth 8, {n_tiles} tiles")
    assert float(ref[..., 1].max()) > 0
    assert werr < 1e-5 and rel_mse < 1e-6 and frac > 0.999
    assert abs(st["rays_extend"] - int(cnt[0])) <= 1e-4 * int(cnt[0])
    ctx.upload(T.scenes.shadows(resolution=32)[0])          # drop the 1.7 GB scene from the shared context


# ---------------------------------------------------------------- SPPM on the named scenes
def _sppm_pair(T, ctx, scene, camera, r0, depth, iters, photons, seed=11, oracle_iters=None):
    flat = ctx.upload(scene)
    osc = oracle_lib.OracleScene(flat)
    cam, fd = camera.pod(), camera.film.desc()
    h, w = camera.film.pixels.shape[:2]
    gpu = np.zeros((h, w, 3), np.float32)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_sppm(ctx.h, C.byref(cam), C.byref(fd), r0, depth, iters, photons, 0, C.c_uint64(seed),
                                        C.cast(None, T._lib.SPPM_CB), None, T._lib.ptr(gpu)))
    st = ctx.stats()
    ref = np.zeros_like(gpu)
    cnt = osc.render_sppm(cam, fd, r0, depth, oracle_iters or iters, photons, seed, ref)
    return gpu, ref, st, cnt


def _rgb_report(gpu, ref, label):
    g4 = np.concatenate([gpu, np.ones_like(gpu[..., :1])], -1)
    r4 = np.concatenate([ref, np.ones_like(ref[..., :1])], -1)
    return image_report(g4, r4, label)[:2]


needs_ply = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "caustic-glass.ply")),
                               reason="asset missing")


@needs_ply
def test_c4_caustic_moving_two_lights(T, ctx):
    """docs/code/caustic_moving.jl: a PointLight AND a SpotLight.  Camera pass: uniform_sample_one_light picks light
    ceil(u * 2) (sppm.jl:503-517); photon pass: sample_discrete over the power distribution (sampling.jl:32-41).  256^2
    with the script's photons-per-pixel ratio, 3 iterations, same iteration count on both sides."""
    scene, camera, kw = T.scenes.caustic_moving(resolution=256)
    assert len(scene.lights) == 2
    photons = 1_250_000 // 16
    gpu, ref, st, cnt = _sppm_pair(T, ctx, scene, camera, kw["initial_search_radius"], kw["max_depth"], 3, photons)
    rel_mse, frac = _rgb_report(gpu, ref, "sppm/caustic-moving (2 lights) 256^2 x 3 it")
    # per-scene tolerance as for caustic_glass: sinf / cosf ULPs of the spot light's photon directions move a few photons
    # across triangle edges of the glass mesh (measured below 5e-3)
    assert rel_mse < 2e-2 and frac > 0.85 and float(ref.max()) > 0
    assert abs(st["rays_extend"] - int(cnt[0])) <= 2e-3 * int(cnt[0])
    # both lights really emit: the image differs from either one-light render
    for keep in (0, 1):
        one = T.Scene([scene.lights[keep]], scene.aggregate)
        g1, _, _, _ = _sppm_pair(T, ctx, one, camera, kw["initial_search_radius"], kw["max_depth"], 1, photons, oracle_iters=1)
        assert np.abs(g1 - gpu).max() > 1e-2 * np.abs(gpu).max()


@needs_ply
@pytest.mark.parametrize("depth", [5, 8])
def test_c1_caustic_glass_as_shipped(T, ctx, depth):
    """docs/code/caustic_glass.jl at its own size (256^2, 65 025 photons per iteration), ray depth 5 (the script) and 8
    (README / BASELINE.json configs[0]), 3 iterations."""
    scene, camera, kw = T.scenes.caustic_glass(resolution=256, max_depth=depth)
    gpu, ref, st, cnt = _sppm_pair(T, ctx, scene, camera, kw["initial_search_radius"], depth, 3, -1)
    rel_mse, frac = _rgb_report(gpu, ref, f"sppm/caustic-glass 256^2 depth {depth} x 3 it")
    # per-scene tolerance: relMSE < 2e-2 (measured 1.6e-3 at depth 8); the share of pixels within 0.2 % of
    # the peak drops with depth - every extra specular bounce is another sinf / cosf / sqrtf ULP that can move a photon
    # across a triangle edge (measured 0.85 at depth 8, 0.9 at depth 5)
    assert rel_mse < 2e-2 and frac > (0.85 if depth <= 5 else 0.8) and float(ref.max()) > 0
    assert abs(st["rays_extend"] - int(cnt[0])) <= 2e-3 * int(cnt[0])


# Converged comparison (north_star: "relative MSE against the reference's converged image must fall within a stated
# tolerance ... reported per scene").  The GPU renders n iterations, the oracle 4n; SPPM's error at n iterations is
# dominated by the photon noise of the n-iteration estimate (variance ~ 1/n), so the tolerance is what the ORACLE ITSELF
# shows between n and 4n iterations, plus 25 %: the test computes that reference gap and asserts the GPU's is no larger.
CONVERGED = [
    # scene builder, kwargs, photons, n, absolute cap on relMSE(GPU n vs oracle 4n)
    ("shadows", dict(resolution=128), -1, 16, 0.12),
    ("caustic_glass", dict(resolution=96), 30_000, 12, 0.25),
]


@pytest.mark.parametrize("name,kw,photons,n,cap", CONVERGED)
def test_converged_image_within_stated_tolerance(T, ctx, name, kw, photons, n, cap):
    if name != "shadows" and not os.path.exists(T.scenes.ASSET_PLY):
        pytest.skip("asset missing")
    scene, camera, ikw = getattr(T.scenes, name)(**kw)
    r0, depth = ikw["initial_search_radius"], ikw["max_depth"]
    gpu, conv, _, _ = _sppm_pair(T, ctx, scene, camera, r0, depth, n, photons, seed=21, oracle_iters=4 * n)
    osc = oracle_lib.OracleScene(scene.flatten())
    same = np.zeros_like(gpu)
    osc.render_sppm(camera.pod(), camera.film.desc(), r0, depth, n, photons, 21, same)
    rel = lambda a, b: float(np.mean((a - b) ** 2) / max(1e-12, np.mean(b ** 2)))
    gap_gpu, gap_oracle = rel(gpu, conv), rel(same, conv)
    print(f"converged/{name}: relMSE GPU({n}) vs oracle({4 * n}) = {gap_gpu:.4g}; oracle({n}) vs oracle({4 * n}) = {gap_oracle:.4g}; "
          f"GPU({n}) vs oracle({n}) = {rel(gpu, same):.4g}")
    assert gap_gpu <= 1.25 * gap_oracle + 1e-4, (gap_gpu, gap_oracle)
    assert gap_gpu < cap

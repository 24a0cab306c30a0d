"""Host-side logic that needs no GPU: scene flattening (incl. the nested-BVH splice), the PLY reader, the scene builders
of the BASELINE.json configs, transformation quirks, multi-GPU partition helpers (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_transformation_quirk_q1(T):
    """t1 * t2 multiplies the inverse matrices in the same order (transformations.jl:20-22)."""
    a, b = T.translate([1, 2, 3]), T.scale(2, 2, 2)
    ab = a * b
    assert np.allclose(ab.m, a.m @ b.m) and np.allclose(ab.inv_m, a.inv_m @ b.inv_m)
    assert not np.allclose(ab.inv_m, np.linalg.inv(ab.m))          # hence not a true inverse
    assert ab.inv().inv() == ab


def test_perspective_is_untransposed(T):
    p = T.perspective(90.0, 0.01, 1000.0)
    # scale(inv_tan) * Transformation(Mat4f(...)) with the column-major fill: element [3,2] holds -f*n/(f-n), [2,3] holds 1
    assert abs(p.m[2, 3] - 1.0) < 1e-6 and abs(p.m[3, 2] + 1000 * 0.01 / (1000 - 0.01)) < 1e-6 and p.m[3, 3] == 0


def test_raster_to_camera_matches_survey_emulation(T):
    """for 1024^2, window (-1,1): pixel (1024,1024) -> (0.01998, 0.02002, -0.99999)."""
    scene, camera, _ = T.scenes.shadows(resolution=1024)
    pc = camera.raster_to_camera.point([1024, 1024, 0])
    assert np.allclose(pc, [0.01998, 0.02002, -0.99999], atol=2e-5)
    p0 = camera.raster_to_camera.point([0, 0, 0])
    assert abs(p0[0]) < 1e-4 and abs(p0[2] + 0.99999) < 1e-4


def test_flatten_orders_and_materials(T):
    scene, camera, _ = T.scenes.shadows(resolution=32)
    flat = scene.flatten()
    assert len(flat.prims) == 8 and len(flat.spheres) == 4 and len(flat.tri_vertices) == 4
    assert sorted(flat.prims["original"].tolist()) == list(range(8))
    assert flat.n_materials == 5 and not flat.has_unshaded
    # leaves cover the ordered primitive list exactly once
    meta = flat.nodes["meta"].astype(np.int64)
    leaves = flat.nodes[(meta >> 30) == 3]
    assert int(((leaves["meta"].astype(np.int64)) & 0x3FFFFFFF).sum()) == 8


def test_nested_bvh_splice_equals_reference_recursion(T):
    """Splicing a nested BVHAccel's nodes in place of its leaf must give the same hits as tracing each BVH separately
    and keeping the nearer hit (what the reference's recursive intersect! does)."""
    rng = np.random.default_rng(0)
    mk = lambda c: T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate(c), False), 0.6, 360.0))
    inner = [mk(rng.uniform(-4, 0, 3)) for _ in range(6)]
    outer = [mk(rng.uniform(0, 4, 3)) for _ in range(5)]
    nested = T.Scene([], T.BVHAccel(outer + [T.BVHAccel(inner)]))
    flat_all = T.Scene([], T.BVHAccel(outer + inner))
    o = rng.uniform(-8, 8, (4000, 3)).astype(np.float32)
    d = (rng.uniform(-4, 4, (4000, 3)) - o).astype(np.float32)
    p1, t1, _ = oracle_lib.OracleScene(nested.flatten()).intersect(o, d)
    p2, t2, _ = oracle_lib.OracleScene(flat_all.flatten()).intersect(o, d)
    assert np.array_equal(p1 != 0, p2 != 0) and np.array_equal(t1, t2)
    assert (p1 != 0).sum() > 100


def test_ply_reader(T):
    if not os.path.exists(T.scenes.ASSET_PLY):
        pytest.skip("asset missing")
    meshes, tris = T.load_triangle_mesh(T.scenes.ASSET_PLY)
    m = meshes[0]
    assert m.n_vertices == 44034 and m.n_triangles == 88064 and len(tris) == 88064
    assert m.indices.min() == 1 and m.indices.max() == 44034          # 1-based like model_loader.jl:35
    assert m.normals is not None and np.allclose(np.linalg.norm(m.normals, axis=1), 1, atol=1e-3)
    assert np.array_equal(tris[5].vertices(), m.vertices[m.indices[15:18].astype(np.int64) - 1])


def test_tessellated_triangle_counts(T):
    v, n, idx = T.scenes._uv_sphere((0, 0, 0), 1.0, 266, 264)
    assert len(idx) == 139920
    v, n, idx = T.scenes._heightfield(600)
    assert len(idx) == 720000
    scene, camera, kw = T.scenes.tessellated(cells=20, stacks=10, slices=8, res=(64, 36))
    assert scene.aggregate.n_primitives == 2 * 20 * 20 + 2 * (2 * 8 * 9)
    assert camera.film.pixels.shape == (36, 64, 4)


def test_sppm_default_photon_count(T):
    scene, camera, kw = T.scenes.caustic_glass(resolution=256) if os.path.exists(T.scenes.ASSET_PLY) else T.scenes.shadows(256)
    integ = T.SPPMIntegrator(camera, 0.075, 5, 100, -1)
    assert integ.photons_per_iteration == 255 * 255                  # area(crop_bounds), Q22


def test_partition_helpers(T):
    from trace_jl_b200 import distributed as D
    for world in (1, 2, 3, 8):
        tiles = [D.tile_shard(527, r, world) for r in range(world)]
        assert sorted(sum(tiles, [])) == list(range(527))
        rng = [D.photon_range(1_046_529, r, world) for r in range(world)]
        assert rng[0][0] == 0 and rng[-1][1] == 1_046_529 and all(rng[i][1] == rng[i + 1][0] for i in range(world - 1))
    scene, camera, _ = T.scenes.shadows(resolution=1024)
    assert D.n_sample_tiles(camera.film) == 65 * 65


def test_sppm_storage_order(T):
    """Row-sharded SPPM arrays: raster <-> storage is a bijection onto the non-padding slots, a rank's pixels fill its
    contiguous slice, and world == 1 is the identity (the layout sppm.cu uses for the all-gather of visible points)."""
    from trace_jl_b200 import distributed as D
    for W, H, world in ((7, 5, 1), (7, 5, 2), (4, 9, 4), (3, 10, 8), (151, 151, 3)):
        chunk, n = D.storage_layout(W, H, world)
        assert chunk == -(-H // world) and n == world * chunk * W
        seen = set()
        for y in range(H):
            for x in range(W):
                s = D.raster_to_storage(x, y, W, H, world)
                owner = y % world
                assert owner * chunk * W <= s < (owner + 1) * chunk * W
                assert D.storage_to_raster(s, W, H, world) == (x, y)
                if world == 1:
                    assert s == y * W + x
                seen.add(s)
        assert len(seen) == W * H
        pads = [s for s in range(n) if D.storage_to_raster(s, W, H, world) is None]
        assert len(pads) == n - W * H and not (set(pads) & seen)


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["REPO"]); sys.path.insert(0, os.path.join(os.environ["REPO"], "tests"))
import numpy as np, torch, torch.distributed as dist
import trace_jl_b200 as T, oracle_lib
from trace_jl_b200 import distributed as D
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# Whitted: each rank renders its round-robin share of the 16x16 tiles with the CPU oracle into a private film; the
# all-reduced film must equal the full render (this is the host-side plan whitted.cu / bench.py follow on GPUs).
scene, camera, _ = T.scenes.shadows(resolution=48)
osc = oracle_lib.OracleScene(scene.flatten())
cam, fd = camera.pod(), camera.film.desc()
n_tiles = D.n_sample_tiles(camera.film)
mine = D.tile_shard(n_tiles, rank, world)
film = np.zeros_like(camera.film.pixels)
import ctypes as C
L = oracle_lib.lib()
osc.render_whitted_tiles(cam, fd, 2, 4, 9, film, mine, threads=1)
t = torch.from_numpy(film)
D.allreduce_sum(t)
full = np.zeros_like(film)
osc.render_whitted(cam, fd, 2, 4, 9, full, threads=1)
assert np.allclose(t.numpy(), full, rtol=1e-5, atol=1e-7), np.abs(t.numpy() - full).max()
# SPPM: row-sharded per-pixel arrays in storage order - each rank fills its slice, the all-gather yields the whole image
W_, H_ = 5, 7
chunk, n_slots = D.storage_layout(W_, H_, world)
buf = torch.full((n_slots,), -1.0)
for s_ in range(rank * chunk * W_, (rank + 1) * chunk * W_):
    xy = D.storage_to_raster(s_, W_, H_, world)
    buf[s_] = float(xy[1] * W_ + xy[0]) if xy else -2.0
parts = [torch.empty(chunk * W_) for _ in range(world)]
dist.all_gather(parts, buf[rank * chunk * W_:(rank + 1) * chunk * W_].clone())
whole = torch.cat(parts)
for y_ in range(H_):
    for x_ in range(W_):
        assert whole[D.raster_to_storage(x_, y_, W_, H_, world)].item() == float(y_ * W_ + x_)
# SPPM photon slices partition the iteration
b, e = D.photon_range(1000, rank, world)
# the communicator id travels from rank 0 to everybody (distributed.init_comm); here with a stand-in id, on gloo
cid = D.broadcast_id(lambda: bytes(range(128)), rank)
assert cid == bytes(range(128)), rank
cnt = torch.tensor([float(e - b)])
D.allreduce_sum(cnt)
assert cnt.item() == 1000
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world2_tile_shard_and_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_scale_and_mix_textures_fold_like_the_reference():
    """ScaleTexture / MixTexture (src/textures/basic.jl:12-37) over constant leaves: t1 * t2 and (1 - t) * t1 + t * t2 in
    Float32, evaluated when the material is flattened; the device receives the constant."""
    import trace_jl_b200 as T
    from trace_jl_b200.scene import _tex_rgb, _tex_f
    f32 = np.float32
    a, b = T.ConstantTexture(T.RGBSpectrum(0.8, 0.5, 0.25)), T.ConstantTexture(T.RGBSpectrum(0.5, 0.25, 0.1))
    t = T.ConstantTexture(0.3)
    got = _tex_rgb(T.ScaleTexture(a, b))
    assert got.dtype == np.float32 and np.array_equal(got, (a.value.c * b.value.c).astype(f32))
    mix = _tex_rgb(T.MixTexture(a, b, t))
    want = ((f32(1) - f32(0.3)) * a.value.c + f32(0.3) * b.value.c).astype(f32)
    assert np.array_equal(mix, want)
    # nested, and scalar textures (roughness / sigma / index)
    nested = _tex_rgb(T.ScaleTexture(T.MixTexture(a, b, t), T.ConstantTexture(T.RGBSpectrum(2.0))))
    assert np.array_equal(nested, (want * f32(2)).astype(f32))
    assert _tex_f(T.MixTexture(T.ConstantTexture(0.1), T.ConstantTexture(0.5), T.ConstantTexture(0.25))) == f32(f32(0.75) * f32(0.1) + f32(0.25) * f32(0.5))
    with pytest.raises(TypeError):
        _tex_f(a)
    with pytest.raises(TypeError):
        _tex_rgb(T.MixTexture(a, b, a))
    # the material POD carries the folded value
    m = T.MatteMaterial(T.ScaleTexture(a, b), T.ConstantTexture(0.0))
    assert np.array_equal(np.asarray(m.pod()[1]), got)


def test_grid_hash_magic_is_an_exact_modulo():
    """sppm.cu:grid_hash replaces `h % n_pixels` (sppm.jl:497-501, UInt64) by q = mulhi64(h, magic) >> shift with
    magic = ceil(2^(57 + L) / n), L = ceil(log2 n): exact for every h < 2^57 (cell coordinates are < 2^30, so the xor of
    the three products stays below 2^57).  The same formula in Python integers, against `//` and `%`."""
    import random
    rng = random.Random(5)

    def magic(n):
        lg = 0
        while (1 << lg) < n:
            lg += 1
        k = 57 + lg
        return -(-(1 << k) // n), k

    for n in [65, 100, 127, 128, 129, 96 * 96, 255 * 255, 1023 * 1023, 1024 * 1024, 1920 * 1080, 4096 * 4096, (1 << 31) - 1]:
        m, k = magic(n)
        assert m < (1 << 64) and k >= 64                       # fits the kernel's 64-bit magic and its `>> (k - 64)`
        hs = [0, 1, n - 1, n, n + 1, (1 << 57) - 1, (1 << 57) - n, ((1 << 57) // n) * n, ((1 << 57) // n) * n - 1]
        hs += [rng.getrandbits(57) for _ in range(20000)]
        hs += [((x * 73856093) ^ (y * 19349663) ^ (z * 83492791)) for x, y, z in
               ((rng.randrange(1 << 30), rng.randrange(1 << 30), rng.randrange(1 << 30)) for _ in range(20000))]
        for h in hs:
            if h >= (1 << 57):                                 # (n divides 2^57: outside the hash's range)
                continue
            q = ((h * m) >> 64) >> (k - 64)
            assert q == h // n and h - q * n == h % n

"""torchrun -N check of the library's own NCCL path (trace_comm_init): every rank renders with the communicator, rank 0
also renders alone on a second context, and the results must agree up to the order of float adds.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_multi_gpu.py

Whitted: film_mode 0 (whole film on rank 0) and film_mode 1 (one band per rank, host-buffer call), device and host entry
points.  SPPM: trace_render_sppm over all ranks.  Prints one JSON line on rank 0; exit code 1 on a mismatch."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import trace_jl_b200 as T
    from trace_jl_b200 import distributed as D
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = T.Context(local, stream=stream.cuda_stream)
    D.init_comm(ctx, rank, world)
    info = ctx.comm_info()
    assert info["world"] == world and info["rank"] == rank and info["nccl_version"] > 0, info
    solo = T.Context(local) if rank == 0 else None
    report = {"world": world, "nccl_version": info["nccl_version"]}
    ok = True

    scene, camera, spp, depth = None, None, 4, 5
    scene, camera, _ = T.scenes.tessellated(cells=64, stacks=34, slices=32, res=(480, 270))
    cam, fd = camera.pod(), camera.film.desc()
    h, w = camera.film.pixels.shape[:2]
    ctx.upload(scene)
    want = None
    if rank == 0:
        solo.upload(scene)
        want = np.zeros((h, w, 4), np.float32)
        solo.check(solo.lib.trace_render_whitted(solo.h, C.byref(cam), C.byref(fd), spp, depth, C.c_uint64(5), T._lib.ptr(want)))
    # film_mode 0, device film: the whole image lands on rank 0, other ranks' films stay untouched
    ctx.set_option("film_mode", 0)
    film = torch.full((h, w, 4), 0.25, dtype=torch.float32, device=f"cuda:{local}")
    ctx.check(ctx.lib.trace_render_whitted_device(ctx.h, C.byref(cam), C.byref(fd), spp, depth, C.c_uint64(5), C.c_void_p(film.data_ptr())))
    torch.cuda.synchronize()
    got = film.cpu().numpy()
    if rank == 0:
        report["whitted_mode0_max_rel_err"] = float(np.abs(got - 0.25 - want).max() / np.abs(want).max())
        ok &= np.allclose(got - 0.25, want, rtol=2e-4, atol=1e-5)
    else:
        ok &= bool(np.all(got == 0.25))
    # film_mode 0, host film: rank 0 passes a film, the others NULL
    host = np.zeros((h, w, 4), np.float32)
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), spp, depth, C.c_uint64(5), T._lib.ptr(host) if rank == 0 else None))
    if rank == 0:
        ok &= np.allclose(host, want, rtol=2e-4, atol=1e-5)
    # film_mode 1: every rank's host film receives its band; the bands assembled must be the image
    ctx.set_option("film_mode", 1)
    band = np.zeros((h, w, 4), np.float32)
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam), C.byref(fd), spp, depth, C.c_uint64(5), T._lib.ptr(band)))
    npix = h * w
    chunk = (npix + world - 1) // world
    flat = band.reshape(-1, 4)
    mine = np.zeros(npix, bool)
    mine[rank * chunk:min(npix, (rank + 1) * chunk)] = True
    ok &= bool(np.all(flat[~mine] == 0))                       # nothing outside the rank's band
    gathered = torch.zeros((world, npix, 4), dtype=torch.float32, device=f"cuda:{local}")
    dist.all_gather_into_tensor(gathered, torch.from_numpy(flat).to(f"cuda:{local}"))
    whole = gathered.sum(0).cpu().numpy().reshape(h, w, 4)
    if rank == 0:
        report["whitted_mode1_max_rel_err"] = float(np.abs(whole - want).max() / np.abs(want).max())
        ok &= np.allclose(whole, want, rtol=2e-4, atol=1e-5)
    ctx.set_option("film_mode", 0)

    # A ray queue that overflows on some rank: every rank must take the redo path together (rank-summed flags) and the image
    # must still equal the one-GPU render.  Glass spawns two rays per hit, so with the queues squeezed to 100 % of a batch
    # bounce level 2 overflows (tests/test_gpu_parity.py has the one-GPU form of this case).
    glass = T.GlassMaterial(T.ConstantTexture(T.RGBSpectrum(1.0)), T.ConstantTexture(T.RGBSpectrum(1.0)), T.ConstantTexture(0.0),
                            T.ConstantTexture(0.0), T.ConstantTexture(1.5), True)
    white = T.MatteMaterial(T.ConstantTexture(T.RGBSpectrum(1.0)), T.ConstantTexture(0.0))
    prims = [T.GeometricPrimitive(T.Sphere(T.ShapeCore(T.translate([0.5, 0.5, -2.5]), False), 0.55, 360.0), glass)]
    tris = T.create_triangle_mesh(T.ShapeCore(T.translate([0, 0, -2]), False), 2, [1, 2, 3, 1, 4, 3], 4,
                                  [[-2, -0.2, 2], [-2, -0.2, -3], [3, -0.2, -3], [3, -0.2, 2]], [[0, 1, 0]] * 4)
    prims += [T.GeometricPrimitive(t, white) for t in tris]
    o_scene = T.Scene([T.PointLight(T.translate([-1, 3, 0]), T.RGBSpectrum(25.0))], T.BVHAccel(prims, 1))
    o_film = T.Film([128, 128], T.Bounds2([0, 0], [1, 1]), T.LanczosSincFilter([1, 1], 3.0), 1.0, 1.0, None)
    o_cam = T.PerspectiveCamera(T.look_at([0, 15, 50], [0, 0, -2], [0, 1, 0]), T.Bounds2([-1, -1], [1, 1]), 0, 1, 0, 1e6, 90.0, o_film)
    oc, of = o_cam.pod(), o_film.desc()
    ctx.upload(o_scene)
    ctx.set_option("cap_percent", 100)
    ctx.set_option("batch", 8192)
    ctx.reset_stats()
    got = np.zeros_like(o_film.pixels)
    ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(oc), C.byref(of), 4, 8, C.c_uint64(5), T._lib.ptr(got) if rank == 0 else None))
    n_over = torch.tensor([float(ctx.stats()["queue_overflows"])], device=f"cuda:{local}")
    dist.all_reduce(n_over)
    ctx.set_option("batch", 1 << 26)
    ctx.set_option("cap_percent", 200)
    if rank == 0:
        solo.upload(o_scene)
        want_o = np.zeros_like(got)
        solo.check(solo.lib.trace_render_whitted(solo.h, C.byref(oc), C.byref(of), 4, 8, C.c_uint64(5), T._lib.ptr(want_o)))
        report["overflow_redo_batches_all_ranks"] = int(n_over.item())
        report["overflow_redo_max_rel_err"] = float(np.abs(got - want_o).max() / np.abs(want_o).max())
        ok &= np.allclose(got, want_o, rtol=2e-4, atol=1e-5) and n_over.item() > 0

    # SPPM over all ranks vs one GPU (two scenes: spheres; the two-light caustic scene when the asset is there)
    cases = [("shadows", T.scenes.shadows(resolution=151), 3, -1)]
    if os.path.exists(T.scenes.ASSET_PLY):
        cases.append(("caustic_moving", T.scenes.caustic_moving(resolution=128), 2, 40_000))
    for name, (s_scene, s_cam, kw), iters, photons in cases:
        sc, sf = s_cam.pod(), s_cam.film.desc()
        sh, sw = s_cam.film.pixels.shape[:2]
        ctx.upload(s_scene)
        a = np.zeros((sh, sw, 3), np.float32)
        ctx.check(ctx.lib.trace_render_sppm(ctx.h, C.byref(sc), C.byref(sf), kw["initial_search_radius"], kw["max_depth"], iters, photons, 0,
                                            C.c_uint64(7), C.cast(None, T._lib.SPPM_CB), None, T._lib.ptr(a)))
        # every rank returns the complete image
        imgs = torch.zeros((world,) + a.shape, dtype=torch.float32, device=f"cuda:{local}")
        dist.all_gather_into_tensor(imgs, torch.from_numpy(a).to(f"cuda:{local}"))
        ok &= bool(torch.equal(imgs[0], imgs[rank]))
        if rank == 0:
            solo.upload(s_scene)
            b = np.zeros_like(a)
            solo.check(solo.lib.trace_render_sppm(solo.h, C.byref(sc), C.byref(sf), kw["initial_search_radius"], kw["max_depth"], iters, photons, 0,
                                                  C.c_uint64(7), C.cast(None, T._lib.SPPM_CB), None, T._lib.ptr(b)))
            report[f"sppm_{name}_max_abs_diff"] = float(np.abs(a - b).max())
            report[f"sppm_{name}_image_max"] = float(b.max())
            ok &= np.allclose(a, b, rtol=3e-4, atol=1e-6) and float(b.max()) > 0
    flag = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["ok"] = bool(flag.item() == 1.0)
    if rank == 0:
        print(json.dumps(report))
        solo.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if report["ok"] else 1)


if __name__ == "__main__":
    main()

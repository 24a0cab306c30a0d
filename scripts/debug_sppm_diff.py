"""Debug helper (GPU box): where do the GPU and oracle SPPM images of caustic_glass differ?"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, trace_jl_b200 as T, oracle_lib

ctx = T.Context(0)
scene, camera, kw = T.scenes.caustic_glass(resolution=64)
flat = ctx.upload(scene)
osc = oracle_lib.OracleScene(flat)
cam, fd = camera.pod(), camera.film.desc()
for depth, iters, photons in ((1, 1, 20000), (5, 1, 20000), (5, 2, 20000), (5, 3, 20000)):
    gpu = np.zeros((64, 64, 3), np.float32); ref = np.zeros_like(gpu)
    ctx.reset_stats()
    ctx.check(ctx.lib.trace_render_sppm(ctx.h, C.byref(cam), C.byref(fd), 0.075, depth, iters, photons, 0, C.c_uint64(11),
                                        C.cast(None, T._lib.SPPM_CB), None, T._lib.ptr(gpu)))
    st = ctx.stats()
    cnt = osc.render_sppm(cam, fd, 0.075, depth, iters, photons, 11, ref)
    d = np.abs(gpu - ref).max(-1)
    peak = ref.max()
    nz = d > 0
    print(f"depth {depth} iters {iters}: peak {peak:.4g} mean {ref.mean():.4g} pixels differing {int(nz.sum())} (> 1e-4 peak: {int((d > 1e-4 * peak).sum())}, > 2e-3 peak: {int((d > 2e-3 * peak).sum())})"
          f" max diff {d.max():.4g} rays gpu {st['rays_extend']} {st['rays_shadow']} ref {cnt} deposits {st['sppm_deposits']}")
    idx = np.argsort(d.ravel())[::-1][:6]
    for i in idx:
        y, x = divmod(int(i), 64)
        print("   px", x, y, "gpu", gpu[y, x], "ref", ref[y, x])

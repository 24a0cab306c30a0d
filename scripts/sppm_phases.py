"""Per-phase GPU time of one SPPM iteration; WORLD>1 plays rank 0 of WORLD on one GPU (no exchange: timing only).
python scripts/sppm_phases.py shadows|caustic_glass|caustic_moving [world] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch, trace_jl_b200 as T
from trace_jl_b200 import distributed as D
name = sys.argv[1] if len(sys.argv) > 1 else "shadows"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
scene, camera, kw = getattr(T.scenes, name)()
torch.cuda.set_device(0)
_s = torch.cuda.Stream(device=0)
torch.cuda.set_stream(_s)
ctx = T.Context(0, stream=_s.cuda_stream)
ctx.upload(scene)
ctx.set_option("world", world)
ctx.set_option("rank", 0)
cam, fd = camera.pod(), camera.film.desc()
photons = kw.get("photons_per_iteration", -1)
if photons <= 0:
    photons = int(camera.film.crop_bounds.area())
ctx.check(ctx.lib.trace_sppm_begin(ctx.h, C.byref(cam), C.byref(fd), kw["initial_search_radius"], kw["max_depth"], photons, C.c_uint64(1)))
names = ["camera", "grid", "photon", "update"]
tot = dict.fromkeys(names, 0.0)
b, e = D.photon_range(photons, 0, world)
for it in range(1, iters + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    ctx.check(ctx.lib.trace_sppm_camera_pass(ctx.h, it))
    if world == 1:
        ev[1].record()          # (world 1: the grid is built inside camera_pass; reported together)
    else:
        ev[1].record()
        ctx.check(ctx.lib.trace_sppm_build_grid(ctx.h))
    ev[2].record()
    ctx.check(ctx.lib.trace_sppm_photon_pass(ctx.h, it, b, e))
    ev[3].record()
    ctx.check(ctx.lib.trace_sppm_update(ctx.h))
    ev[4].record()
    torch.cuda.synchronize()
    if it > 2:
        for k, nm in enumerate(names):
            tot[nm] += ev[k].elapsed_time(ev[k + 1])
print(f"{name} world {world}: " + "  ".join(f"{nm} {tot[nm] / iters:.3f} ms" for nm in names) + f"  total {sum(tot.values()) / iters:.3f} ms")
ctx.check(ctx.lib.trace_sppm_end(ctx.h))

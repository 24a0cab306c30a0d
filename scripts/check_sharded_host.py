"""torchrun -N check: ShardedWhittedRenderer (host film, upload overlapped, reduce) == trace_render_whitted on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch, torch.distributed as dist
import trace_jl_b200 as T
from trace_jl_b200 import distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
st = torch.cuda.Stream(device=local)
torch.cuda.set_stream(st)
ctx = T.Context(local, stream=st.cuda_stream)
scene, camera, _ = T.scenes.tessellated(cells=48, stacks=26, slices=24, res=(320, 180))
H, W = camera.film.pixels.shape[:2]
init = np.random.default_rng(3).random((H, W, 4), dtype=np.float32)
host = torch.from_numpy(init.copy()).pin_memory() if rank == 0 else None
r = D.ShardedWhittedRenderer(ctx, scene, camera, rank, world)
for seed in (5, 6):
    r.render(host, 4, 5, seed)
if rank == 0:
    solo = T.Context(local, stream=st.cuda_stream)
    solo.upload(scene)
    ref = init.copy()
    cam, fd = camera.pod(), camera.film.desc()
    for seed in (5, 6):
        solo.check(solo.lib.trace_render_whitted(solo.h, C.byref(cam), C.byref(fd), 4, 5, C.c_uint64(seed), T._lib.ptr(ref)))
    got = host.numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print("sharded host film vs single GPU: max rel err", float(err), "OK" if err < 1e-5 else "MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()

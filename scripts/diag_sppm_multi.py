"""Diagnostic: the bench's SPPM sequence (session with per-kernel timing, then repeated trace_render_sppm with changing seeds)
over all ranks, with progress lines, no torch collectives after the communicator id broadcast.
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      scripts/diag_sppm_multi.py --pipeline 4 [--workload sppm-shadows-1024] [--renders 4]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pipeline", type=int, default=4)
    ap.add_argument("--workload", default="sppm-shadows-1024")
    ap.add_argument("--renders", type=int, default=4)
    ap.add_argument("--skip-session", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import trace_jl_b200 as T
    import bench
    from trace_jl_b200 import distributed as D
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = T.Context(local, stream=stream.cuda_stream)
    D.init_comm(ctx, rank, world)
    ctx.set_option("sppm_pipeline", args.pipeline)

    def say(msg):
        print(f"[rank {rank}] {msg}", flush=True)

    scene, camera, p = bench.build_sppm_scene(T, args.workload)
    ctx.upload(scene)
    h, w = camera.film.pixels.shape[:2]
    cam, fd = camera.pod(), camera.film.desc()
    if not args.skip_session:
        sess = D.SPPMSession(ctx, scene, camera, p["r0"], p["max_depth"], p["photons"], 0x5EED0001)
        sess.step(3); ctx.synchronize(); say("session: 3 steps ok")
        for _ in range(10):
            sess.step(1)
        ctx.synchronize(); say("session: 10 single steps ok")
        ctx.set_option("time_kernels", 1)
        sess.step(1); ctx.synchronize(); ctx.reset_stats()
        for _ in range(8):
            sess.step(1)
        img = sess.image(); say(f"session: timed steps + image ok, mean {img.mean():.4f}")
        ctx.set_option("time_kernels", 0)
        sess.close()
        ctx.set_option("count_nodes", 1)
        s2 = D.SPPMSession(ctx, scene, camera, p["r0"], p["max_depth"], p["photons"], 0x5EED0001)
        ctx.reset_stats()
        s2.step(1)
        ctx.synchronize()
        st = ctx.stats()
        s2.close()
        ctx.set_option("count_nodes", 0)
        say(f"count_nodes session ok: {st['nodes_visited']} nodes")
    rgb = np.zeros((h, w, 3), np.float32)
    for i in range(args.renders):
        ctx.check(ctx.lib.trace_render_sppm(ctx.h, C.byref(cam), C.byref(fd), p["r0"], p["max_depth"], 10, p["photons"], 0,
                                            C.c_uint64(0x5EED0001 + i), C.cast(None, T._lib.SPPM_CB), None, T._lib.ptr(rgb)))
        say(f"render {i} ok, mean {rgb.mean():.4f}")
    ctx.close()
    say("done")


if __name__ == "__main__":
    main()

"""Per-class timeline of a few pipelined SPPM iterations on every rank (option time_kernels + TRACE_CUDA_TIMELINE), with the
library's collectives as their own class.  Rank 0 prints, per class, the summed launch time and the busy time (union of
the launches' intervals), and the collectives one by one.
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
      scripts/sppm_timeline_multi.py [--workload sppm-caustic-moving] [--iters 8] [--pipeline 4]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sppm-caustic-moving")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--pipeline", type=int, default=4)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    path = f"/tmp/trace_timeline_rank{rank}.txt"
    os.environ["TRACE_CUDA_TIMELINE"] = path
    import torch
    import torch.distributed as dist
    import trace_jl_b200 as T
    import bench
    from trace_jl_b200 import distributed as D
    from trace_jl_b200 import _lib
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = T.Context(local, stream=stream.cuda_stream)
    D.init_comm(ctx, rank, world)
    ctx.set_option("sppm_pipeline", args.pipeline)
    scene, camera, p = bench.build_sppm_scene(T, args.workload)
    sess = D.SPPMSession(ctx, scene, camera, p["r0"], p["max_depth"], p["photons"], 0x5EED0001)
    sess.step(6)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sess.step(args.iters); e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / args.iters
    ctx.set_option("time_kernels", 1)
    if os.path.exists(path):
        os.remove(path)
    e0.record(); sess.step(args.iters); e1.record(); torch.cuda.synchronize()
    timed = e0.elapsed_time(e1) / args.iters
    sess.image()                         # collects the events and writes the timeline
    sess.close()
    if rank == 0:
        rows = [l.split() for l in open(path) if not l.startswith("#")]
        ev = [(int(l), int(k), float(a), float(b)) for l, k, a, b in rows]
        span = max(b for _, _, _, b in ev) - min(a for _, _, a, _ in ev)
        print(f"{args.workload} world {world} pipeline {args.pipeline}: {plain:.3f} ms/iteration ({timed:.3f} with per-launch events); "
              f"timeline span {span:.3f} ms for {args.iters} iterations")
        for k, name in enumerate(_lib.KIND_NAMES):
            iv = sorted((a, b) for _, kk, a, b in ev if kk == k)
            if not iv:
                continue
            total = sum(b - a for a, b in iv)
            busy, cur_a, cur_b = 0.0, iv[0][0], iv[0][1]
            for a, b in iv[1:]:
                if a > cur_b:
                    busy += cur_b - cur_a; cur_a, cur_b = a, b
                else:
                    cur_b = max(cur_b, b)
            busy += cur_b - cur_a
            print(f"  {name:9s} launches {len(iv):4d}  summed {total / args.iters:7.3f} ms/it  busy {busy / args.iters:7.3f} ms/it")
        comm = sorted((a, b, l) for l, kk, a, b in ev if kk == 8)
        print("  collectives [start, end] ms:", "  ".join(f"[{a:.2f},{b:.2f}]" for a, b, _ in comm[:24]))
        upd = sorted((a, b) for _, kk, a, b in ev if kk == 7)
        print("  updates     [start, end] ms:", "  ".join(f"[{a:.2f},{b:.2f}]" for a, b in upd[:12]))
    ctx.close()


if __name__ == "__main__":
    main()

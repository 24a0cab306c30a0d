#!/usr/bin/env python
"""A/B of traversal-loop variants on one GPU, inside one process (same box, same scene upload):
  python scripts/trav_ab.py [--workload tess-1M] [--variants leaf_wait=0 leaf_wait=8 ...] [--steps 5]
Each variant is a comma-separated list of trace_set_option key=value pairs.  For each: bit-equality of 1M closest-hit /
any-hit queries against the first variant, then Whitted ms/step as benchmarked (lanes, graph) and the per-launch
extend / shadow split of a serial pass (lanes = 1, CUDA events around every traversal launch).
Writes JSON lines to gpurun_out/trav_ab.jsonl."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="tess-1M")
    ap.add_argument("--variants", nargs="+", default=["leaf_wait=0", "leaf_wait=4", "leaf_wait=8", "leaf_wait=16", "leaf_wait=32"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--lanes", type=int, default=12)
    ap.add_argument("--builder", default="reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trav_ab.jsonl"))
    args = ap.parse_args()
    import torch
    import trace_jl_b200 as T
    import bench
    from trace_jl_b200 import distributed as D

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = T.Context(0, stream=stream.cuda_stream)
    t0 = time.time()
    scene, camera, spp, depth = bench.build_scene(T, args.workload, args.builder)
    flat = ctx.upload(scene)
    print(f"scene built + uploaded in {time.time() - t0:.1f} s", flush=True)
    H, W = camera.film.pixels.shape[:2]
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
    rng = np.random.default_rng(5)
    n = 1_000_000
    lo, hi = np.array(flat.nodes[0]["bmin"]), np.array(flat.nodes[0]["bmax"])
    o = rng.uniform(lo - 5, hi + 5, (n, 3)).astype(np.float32)
    d = (rng.uniform(lo, hi, (n, 3)) - o).astype(np.float32)
    base = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = open(args.out, "a")
    for var in args.variants:
        opts = dict(kv.split("=") for kv in var.split(",") if kv)
        lanes = int(opts.pop("lanes", args.lanes))
        world = int(opts.pop("world", 1))
        for k, v in opts.items():
            ctx.set_option(k, int(v))
        prim, t, b = ctx.intersect(o, d)
        occ = ctx.occluded(o, d)
        if base is None:
            base = (prim, t, b, occ)
        same = bool(np.array_equal(prim, base[0]) and np.array_equal(t.view(np.uint32), base[1].view(np.uint32)) and
                    np.array_equal(b.view(np.uint32), base[2].view(np.uint32)) and np.array_equal(occ, base[3]))

        def step(i):
            film.zero_()
            D.render_whitted_sharded(ctx, scene, camera, spp, depth, 1000 + i, film, 0, world, reduce=False)

        ctx.set_option("lanes", lanes)
        ctx.set_option("time_kernels", 0)
        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        ctx.reset_stats()
        e0.record()
        for i in range(args.steps):
            step(2 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        st = ctx.stats()
        rays = (st["rays_extend"] + st["rays_shadow"]) / args.steps
        checksum = float(film.double().sum().item())
        # serial pass
        ctx.set_option("lanes", 1)
        ctx.set_option("time_kernels", 1)
        step(100)
        torch.cuda.synchronize()
        ctx.reset_stats()
        e0.record()
        for i in range(3):
            step(101 + i)
        e1.record()
        torch.cuda.synchronize()
        s2 = ctx.stats()
        rec = {"variant": var, "workload": args.workload, "same_hits_as_first": same, "ms_per_step": ms, "Mrays_per_s": rays / ms / 1e3,
               "serial_ms_per_step": e0.elapsed_time(e1) / 3, "extend_ms_per_step": s2["ms_extend"] / 3,
               "shadow_ms_per_step": s2["ms_shadow"] / 3, "film_checksum": checksum, "hit_fraction": float((prim != 0).mean())}
        ctx.set_option("time_kernels", 0)
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        out.flush()
    ctx.close()


if __name__ == "__main__":
    main()

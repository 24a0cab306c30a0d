#!/usr/bin/env python
"""SPPM A/B on one GPU: iterations/s of the named scenes for each option variant (comma-separated key=value lists),
plus the image difference against the first variant after the same iterations.
  python scripts/sppm_bench.py --variants sppm_path=0 sppm_path=1 [--workloads sppm-shadows-1024,sppm-caustic-moving]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", nargs="+", default=["sppm_path=0", "sppm_path=1"])
    ap.add_argument("--workloads", default="sppm-shadows-1024,sppm-caustic-moving,sppm-caustic-glass-d8")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import torch
    import bench
    import trace_jl_b200 as T
    from trace_jl_b200 import distributed as D
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = T.Context(0, stream=stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for wl in args.workloads.split(","):
        scene, camera, p = bench.build_sppm_scene(T, wl)
        base = None
        for var in args.variants:
            for kv in var.split(","):
                k, v = kv.split("=")
                ctx.set_option(k, int(v))
            sess = D.SPPMSession(ctx, scene, camera, p["r0"], p["max_depth"], p["photons"], 0x5EED0001)
            sess.step(3)
            img = sess.image()
            torch.cuda.synchronize()
            ctx.reset_stats()
            e0.record()
            sess.step(args.iters)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            st = ctx.stats()
            sess.close()
            if base is None:
                base = img
            rel = float(np.mean((img - base) ** 2) / max(1e-12, np.mean(base ** 2)))
            print(json.dumps({"workload": wl, "variant": var, "it_per_s": 1e3 / ms, "ms_per_iteration": ms, "relMSE_vs_first_after_3_it": rel,
                              "max_abs_diff": float(np.abs(img - base).max()), "rays_per_it": (st["rays_extend"] + st["rays_shadow"]) / args.iters,
                              "launches_per_it": st["kernel_launches"] / args.iters}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — headline benchmark of the ray-tracing hot path (BASELINE.json: "Mrays/sec (closest-hit + shadow)").

A step is one pass of the hot path over one batch of synthetic input: one full Whitted render (depth 5, 16 spp,
1920x1080) of the synthetic ~1M-triangle "tess-1M" scene (BASELINE.json configs[2]) — the
configuration the north_star's throughput target is quoted on.  value = (rays through closest-hit traversal + rays
through any-hit traversal) / time, whole job, scene and film resident in HBM.  e2e = the same metric through the
host-buffer C ABI call (trace_render_whitted: film H2D + D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload tess-1M|tess-small]
N > 1: launched by torch.distributed.run, one rank per GPU, scene replicated, 16x16 sample tiles dealt round-robin,
one NCCL reduce of the film per step (strong scaling: the image is fixed).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene kwargs, spp, depth)
    "tess-1M": (dict(cells=600, stacks=266, slices=264, res=(1920, 1080), window=((-50.0, -28.125), (50.0, 28.125))), 16, 5),
    # BASELINE.json configs[4] (C5): 10 002 224 triangles, 4096^2 x 64 spp, depth 8 - meant for 8 GPUs
    "tess-10M": (dict(cells=1900, stacks=835, slices=834, res=(4096, 4096), window=((-50.0, -50.0), (50.0, 50.0))), 64, 8),
    "tess-small": (dict(cells=64, stacks=34, slices=32, res=(480, 270), window=((-50.0, -28.125), (50.0, 28.125))), 4, 5),
}
METRIC = "Mrays/sec (closest-hit + shadow)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line), sampled in-process through
    NVML every few ms - the timed region is short (K x 30 ms) and `nvidia-smi -lms` needs longer than that to start.
    Falls back to one nvidia-smi query when NVML is not importable."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.running = False
        self.th = None
        self.nvml = None

    def _handle(self):
        import pynvml
        pynvml.nvmlInit()
        self.nvml = pynvml
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            return pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _loop(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while self.running:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.h))
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            self.h = self._handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.running = True
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception:
            self.nvml = None

    def stop(self):
        if self.th:
            self.running = False
            self.th.join(timeout=2)
        if not self.sm:                      # no NVML: one nvidia-smi sample right after the timed region
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml, 5 ms period, during the timed region"}


def build_scene(T, workload, builder="reference"):
    kw, spp, depth = WORKLOADS[workload]
    scene, camera, _ = T.scenes.tessellated(**kw, builder=builder)
    return scene, camera, spp, depth


def calibrate_tiles(osc, cam, fd, spp, depth, film, cores, total_tiles, target_s):
    """Number of 16x16 tiles (strided over the image) the CPU restatement renders in about `target_s` seconds."""
    tiles = min(total_tiles, max(cores * 4, 64))
    for _ in range(5):
        t0 = time.time()
        osc.render_whitted(cam, fd, spp, depth, 1, film, max_tiles=tiles, threads=cores)
        dt = time.time() - t0
        if tiles >= total_tiles or dt >= 0.6 * target_s:
            break
        tiles = int(min(total_tiles, tiles * min(16.0, max(1.5, target_s / max(dt, 1e-3)))))
    return tiles


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the C++ restatement, oracle/ — Julia is
    not available) with all host threads, on a bounded sample (a strided subset of the 16x16 tiles) of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import trace_jl_b200 as T
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    scene, camera, spp, depth = build_scene(T, args.workload)
    osc = oracle_lib.OracleScene(scene.flatten())
    cam, fd = camera.pod(), camera.film.desc()
    cores = os.cpu_count() or 1
    film = np.zeros_like(camera.film.pixels)
    from trace_jl_b200.distributed import n_sample_tiles
    total_tiles = n_sample_tiles(camera.film)
    # bounded sample per step: ~ref_seconds of CPU work, and the whole run within ~2 minutes
    target = min(args.ref_seconds, 120.0 / max(1, args.steps + args.warmup))
    tiles = calibrate_tiles(osc, cam, fd, spp, depth, film, cores, total_tiles, target)
    times, rays = [], []
    for i in range(args.warmup + args.steps):
        film[:] = 0
        t0 = time.time(); cnt = osc.render_whitted(cam, fd, spp, depth, 1 + i, film, max_tiles=tiles, threads=cores); dt = time.time() - t0
        if i >= args.warmup:
            times.append(dt); rays.append(int(cnt[0]) + int(cnt[1]))
    value = sum(rays) / sum(times) / 1e6
    sample = f"{tiles} of {total_tiles} 16x16 sample tiles (strided over the image) per step, {spp} spp, depth {depth}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"whitted-{args.workload}", "spp": spp, "max_depth": depth,
                   "resolution": list(WORKLOADS[args.workload][0]["res"]), "triangles": int(scene.aggregate.n_primitives)},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated reference (C++ oracle), not Trace.jl itself: no Julia in this image"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="tess-1M", choices=sorted(WORKLOADS))
    ap.add_argument("--slab", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--persist", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=12)
    ap.add_argument("--walk", type=int, default=1, help="traversal loop: 1 pair nodes (default), 0 one node per step (the reference loop)")
    ap.add_argument("--builder", default="reference", choices=["reference", "sah"],
                    help="BVH build: the reference's split logic (default, bit-identical tree) or the opt-in conventional SAH")
    ap.add_argument("--graph", type=int, default=1, help="replay the render as one CUDA graph (0: direct launches)")
    ap.add_argument("--ref-seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sppm", action="store_true")
    ap.add_argument("--no-optin", action="store_true", help="skip the extra measurement on the opt-in SAH tree")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import trace_jl_b200 as T
    from trace_jl_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    # one non-default torch stream for everything: the library's launches, torch's fills, NCCL's ordering and the CUDA
    # events below all see the same stream (torch's default stream has handle 0, which the library takes as "make your own")
    work_stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(work_stream)
    ctx = T.Context(local, stream=work_stream.cuda_stream)
    ctx.set_option("slab", args.slab)
    ctx.set_option("persist", args.persist)
    ctx.set_option("walk", args.walk)
    ctx.set_option("lanes", args.lanes)
    ctx.set_option("graph", args.graph)
    if args.batch:
        ctx.set_option("batch", args.batch)

    scene, camera, spp, depth = build_scene(T, args.workload, args.builder)
    flat = ctx.upload(scene)
    H, W = camera.film.pixels.shape[:2]
    film_dev = torch.zeros((H, W, 4), dtype=torch.float32, device=f"cuda:{local}")
    film_bytes = film_dev.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    parts = {"render": 0.0, "reduce": 0.0, "n": 0}

    def step(i, timed=False):
        film_dev.zero_()
        if timed:
            ev[0].record()
        D.render_whitted_sharded(ctx, scene, camera, spp, depth, 1000 + i, film_dev, rank, world, reduce=False)
        if timed:
            ev[1].record()
        if world > 1:
            dist.reduce(film_dev, dst=0, op=dist.ReduceOp.SUM)
        if timed:
            ev[2].record()
            torch.cuda.synchronize()
            parts["render"] += ev[0].elapsed_time(ev[1]); parts["reduce"] += ev[1].elapsed_time(ev[2]); parts["n"] += 1

    # ---- device-resident timing: W warm-up steps, then exactly K steps between barriers; max over ranks
    ctx.set_option("time_kernels", 0)
    for i in range(args.warmup):
        step(i)
    barrier()
    ctx.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
    st = ctx.stats()
    clocks = sampler.stop() if rank == 0 else None
    rays = torch.tensor([float(st["rays_extend"]), float(st["rays_shadow"]), float(st["kernel_launches"])], device=f"cuda:{local}",
                        dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    total_rays = float(rays[0] + rays[1])
    value = total_rays / (ms_total * 1e-3) / 1e6
    launches = int(rays[2].item())

    # ---- roofline of the dominant kernel (closest-hit `extend`): algorithmic bytes / CUDA-event launch time.
    # In the timed region above up to `lanes` sub-batches run concurrently on side streams, so per-kernel event times
    # overlap; the per-launch durations are therefore taken in a second timed pass of the same K steps with lanes = 1
    # (kernels back to back on one stream), CUDA events around every extend / shadow launch.
    ctx.set_option("lanes", 1)
    ctx.set_option("time_kernels", 1)
    step(5_000)
    barrier()
    ctx.reset_stats()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for i in range(args.steps):
        step(5_001 + i)
    r1.record()
    barrier()
    serial_ms = r0.elapsed_time(r1)
    st = ctx.stats()
    ext_ms, ext_n = st["ms_extend"], max(1, st["extend_launches"])
    ctx.set_option("time_kernels", 0)
    ctx.set_option("count_nodes", 1)
    ctx.reset_stats()
    step(10_000)                                    # instrumented pass (not timed): nodes / primitives per ray
    torch.cuda.synchronize()
    sc = ctx.stats()
    ctx.set_option("count_nodes", 0)
    ctx.set_option("lanes", args.lanes)
    n_rays_cnt = max(1, sc["rays_extend"] + sc["rays_shadow"])
    nodes_per_ray = sc["nodes_visited"] / n_rays_cnt
    prims_per_ray = sc["prims_tested"] / n_rays_cnt
    bytes_per_ray = 48.0 + 32.0 * nodes_per_ray + 48.0 * prims_per_ray
    ext_rays = st["rays_extend"]
    achieved = (ext_rays * bytes_per_ray) / max(1e-9, ext_ms * 1e-3) / 1e9
    peak, peak_kind = load_peaks()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "k_wh_extend", "peak_source": f"of {peak_kind}", "avg_launch_ms": ext_ms / ext_n,
                "launches_timed": int(ext_n), "algorithmic_bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
                "prims_per_ray": prims_per_ray, "extend_Mrays_per_s": ext_rays / max(1e-9, ext_ms * 1e-3) / 1e6,
                "extend_share_of_step": ext_ms / max(1e-9, serial_ms), "serial_pass_ms_per_step": serial_ms / args.steps,
                "shadow_avg_launch_ms": st["ms_shadow"] / max(1, st["shadow_launches"]),
                "shadow_share_of_step": st["ms_shadow"] / max(1e-9, serial_ms),
                "note": "traversal is L1/L2-latency bound: algorithmic bytes are mostly cache hits, not HBM traffic"}
    if os.path.exists(os.path.join(ROOT, "profiles", "extend_traffic.json")):
        try:
            roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "extend_traffic.json"))).get(args.workload)
        except Exception:
            pass
    try:
        # what actually crosses the HBM interface (ncu dram bytes per launch) over the same live launch time: the gap to
        # `achieved` is the share of the algorithmic bytes served by L1/L2
        if roofline.get("traffic"):
            roofline["dram_achieved_GBps"] = float(roofline["traffic"]) / (roofline["avg_launch_ms"] * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_achieved_GBps"] / float(roofline["peak"])
    except Exception:
        pass

    # ---- per-rank breakdown of a step (separate, untimed-for-the-metric pass): render vs film reduce
    breakdown = None
    if world > 1:
        for i in range(3):
            step(20_000 + i, timed=True)
        bd = torch.tensor([parts["render"] / parts["n"], parts["reduce"] / parts["n"]], device=f"cuda:{local}")
        bmax, bmin = bd.clone(), bd.clone()
        dist.all_reduce(bmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(bmin, op=dist.ReduceOp.MIN)
        breakdown = {"render_ms_max_over_ranks": float(bmax[0]), "render_ms_min_over_ranks": float(bmin[0]),
                     "reduce_ms_max_over_ranks": float(bmax[1]), "reduce_ms_min_over_ranks": float(bmin[1])}

    # ---- e2e: the reference-facing C ABI call with HOST buffers (pinned), H2D + D2H inside the timed region
    host_film = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    cam_pod, fd = camera.pod(), camera.film.desc()

    sharded_host = D.ShardedWhittedRenderer(ctx, scene, camera, rank, world) if world > 1 else None

    def e2e_step(i):
        if world == 1:
            ctx.check(ctx.lib.trace_render_whitted(ctx.h, C.byref(cam_pod), C.byref(fd), spp, depth, C.c_uint64(2000 + i),
                                                   C.c_void_p(host_film.data_ptr())))
        else:
            sharded_host.render(host_film if rank == 0 else None, spp, depth, 2000 + i)

    e2e_step(0)
    barrier()
    ctx.reset_stats()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        e2e_step(1 + i)            # (the film accumulates over the steps, as a reference film does over renders)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    st2 = ctx.stats()
    ms2 = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=f"cuda:{local}")
    rays2 = torch.tensor([float(st2["rays_extend"] + st2["rays_shadow"])], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays2, op=dist.ReduceOp.SUM)
    e2e_value = float(rays2.item()) / (float(ms2.item()) * 1e-3) / 1e6
    e2e = {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(film_bytes + C.sizeof(cam_pod) + C.sizeof(fd)),
           "d2h_bytes_per_step": int(film_bytes), "ms_per_step": float(ms2.item()) / args.steps}

    # ---- the same workload on the OPT-IN tree: BVHAccel(..., builder="sah") - same hits (t bit-identical,
    # primitives equal except ties of equal t: tests/test_gpu_parity.py::test_optin_sah_tree_on_gpu), fewer box tests.
    # Reported next to the headline, which stays on the reference's own tree.
    optin = None
    if args.builder == "reference" and not args.no_optin:
        scene2, camera2, _, _ = build_scene(T, args.workload, "sah")
        ctx.upload(scene2)

        def step2(i):
            film_dev.zero_()
            D.render_whitted_sharded(ctx, scene2, camera2, spp, depth, 3000 + i, film_dev, rank, world, reduce=False)
            if world > 1:
                dist.reduce(film_dev, dst=0, op=dist.ReduceOp.SUM)

        for i in range(args.warmup):
            step2(i)
        barrier()
        ctx.reset_stats()
        barrier()
        e0.record()
        for i in range(args.steps):
            step2(args.warmup + i)
        e1.record()
        barrier()
        ms4 = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
        st4 = ctx.stats()
        rays4 = torch.tensor([float(st4["rays_extend"] + st4["rays_shadow"])], device=f"cuda:{local}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
            dist.all_reduce(rays4, op=dist.ReduceOp.SUM)
        optin = {"bvh_builder": "opt-in conventional binned SAH (trace_bvh_build_sah)", "value": float(rays4.item()) / (float(ms4.item()) * 1e-3) / 1e6,
                 "unit": "Mrays/s", "ms_per_step": float(ms4.item()) / args.steps, "bvh_nodes": int(len(scene2.flatten().nodes))}
        ctx.upload(scene)

    # ---- secondary metric: SPPM iterations/s, photons sharded over the ranks (camera pass by image rows):
    # configs[1] docs/code/spheres.jl ("shadows", 1024^2, depth 5) and configs[3] docs/code/caustic_moving.jl (one frame's
    # scene, 1024^2, 1.25 M photons per iteration, depth 5)
    sppm = None
    if not args.no_sppm:
        sppm = []
        for name, make in (("sppm-shadows-1024", lambda: T.scenes.shadows(resolution=1024)),
                           ("sppm-caustic-moving-1024", lambda: T.scenes.caustic_moving())):
            s_scene, s_cam, kw = make()
            sess = D.SPPMSession(ctx, s_scene, s_cam, kw["initial_search_radius"], kw["max_depth"],
                                 kw.get("photons_per_iteration", -1), 0x5EED0001, rank, world)
            for _ in range(3):
                sess.step()
            barrier()
            n_it = 10
            e0.record()
            for _ in range(n_it):
                sess.step()
            e1.record()
            barrier()
            ms3 = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
            if world > 1:
                dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
            sppm.append({"metric": "SPPM iterations/sec", "value": n_it / (float(ms3.item()) * 1e-3), "unit": "it/s",
                         "config": {"workload": name, "photons_per_iteration": sess.photons, "max_depth": kw["max_depth"],
                                    "primitives": int(s_scene.aggregate.n_primitives)}})
            sess.close()
        if world > 1:
            # self-check of the sharded SPPM path on the real collectives: a small render over all ranks against the same
            # render on rank 0 alone (second context, world 1); the sharding must not change the image beyond the order
            # of the float flux atomics
            c_scene, c_cam, ckw = T.scenes.shadows(resolution=160)
            sess = D.SPPMSession(ctx, c_scene, c_cam, ckw["initial_search_radius"], ckw["max_depth"], -1, 0x5EED0001, rank, world)
            for _ in range(4):
                sess.step()
            img_sharded = sess.image()
            sess.close()
            check = None
            if rank == 0:
                solo = T.Context(local, stream=work_stream.cuda_stream)
                s1 = D.SPPMSession(solo, c_scene, c_cam, ckw["initial_search_radius"], ckw["max_depth"], -1, 0x5EED0001, 0, 1)
                for _ in range(4):
                    s1.step()
                img_solo = s1.image()
                s1.close()
                solo.close()
                err = float(np.abs(img_sharded - img_solo).max())
                check = {"workload": "sppm-shadows-160, 4 iterations", "max_abs_diff_vs_one_gpu": err, "image_max": float(img_solo.max()),
                         "ok": bool(np.allclose(img_sharded, img_solo, rtol=3e-4, atol=1e-6))}
            sppm.append({"sharded_vs_single_gpu_check": check})
            ctx.set_option("world", world)
            ctx.set_option("rank", rank)
        ctx.upload(scene)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample of the same workload
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        osc = oracle_lib.OracleScene(flat)
        cores = os.cpu_count() or 1
        tmp = np.zeros_like(camera.film.pixels)
        total_tiles = D.n_sample_tiles(camera.film)
        tiles2 = calibrate_tiles(osc, cam_pod, fd, spp, depth, tmp, cores, total_tiles, 12.0)
        t0 = time.time(); cnt = osc.render_whitted(cam_pod, fd, spp, depth, 1, tmp, max_tiles=tiles2, threads=cores); dt = time.time() - t0
        cpu_baseline = {"value": (int(cnt[0]) + int(cnt[1])) / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                        "sample": f"{tiles2} of {total_tiles} 16x16 sample tiles (strided over the image), {spp} spp, depth {depth}, {dt:.1f} s",
                        "note": "restated reference (C++ oracle, std::thread over tiles), not Trace.jl itself"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"whitted-{args.workload}", "spp": spp, "max_depth": depth,
                          "resolution": list(WORKLOADS[args.workload][0]["res"]), "triangles": int(scene.aggregate.n_primitives),
                          "bvh_nodes": int(len(flat.nodes)),
                          "bvh_builder": {"reference": "reference split logic (src/accel/bvh.jl:87-185), bit-identical tree",
                                          "sah": "opt-in conventional binned SAH (same hits, ties aside)"}[args.builder],
                          "slab_test": {0: "literal (bounds.jl:180-200)", 1: "textbook (not hit-equivalent)", 2: "guarded (literal AND conservative interval; hit-identical, tests/test_gpu_parity.py)"}[args.slab],
                          "parallelism": f"tiles-rr{world}", "lanes": args.lanes, "l2_policy": "inputs larger than L2 (ray queues + BVH > 126 MB per step)",
                          "rays_per_step": total_rays / args.steps},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
               "sppm": sppm, "optin_sah_tree": optin, "breakdown": breakdown}
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

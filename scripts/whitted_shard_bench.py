"""One GPU playing rank 0 of a `world`-GPU Whitted render (tiles k = 0, world, 2*world, ...): the per-rank cost of the
strong-scaling bench without needing `world` GPUs.  usage: whitted_shard_bench.py WORLD [lanes ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trace_jl_b200 as T
from trace_jl_b200 import distributed as D
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp_override = int(os.environ.get("SPP", "0"))
graphs = [int(g) for g in os.environ.get("GRAPH", "0,1").split(",")]
lanes_list = [int(a) for a in sys.argv[2:]] or [12, 6, 4, 2, 1]
_s = torch.cuda.Stream(device=0)
torch.cuda.set_stream(_s)
ctx = T.Context(0, stream=_s.cuda_stream)
kw_b, spp, depth = bench.WORKLOADS["tess-1M"]
scene, camera, _ = T.scenes.tessellated(**kw_b, builder=os.environ.get("BUILDER", "reference"), max_node_primitives=int(os.environ.get("MNP", "1")))
spp = spp_override or spp
depth = int(os.environ.get("DEPTH", depth))
H, W = camera.film.pixels.shape[:2]
film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
if os.environ.get("BATCH"):
    ctx.set_option("batch", int(os.environ["BATCH"]))
ctx.set_option("deal", int(os.environ.get("DEAL", "-2")))
for graph in graphs:
    for lanes in lanes_list:
        ctx.set_option("graph", graph)
        ctx.set_option("lanes", lanes)
        for i in range(3):
            D.render_whitted_sharded(ctx, scene, camera, spp, depth, 7 + i, film, 0, world, reduce=False)
        torch.cuda.synchronize()
        ctx.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n):
            D.render_whitted_sharded(ctx, scene, camera, spp, depth, 100 + i, film, 0, world, reduce=False)
        e1.record()
        torch.cuda.synchronize()
        st = ctx.stats()
        ms = e0.elapsed_time(e1) / n
        rays = (st["rays_extend"] + st["rays_shadow"]) / n
        print(f"builder {os.environ.get('BUILDER', 'reference')}/{os.environ.get('MNP', '1')} depth {depth} spp {spp} world {world} graph {graph} lanes {lanes:2d}: {ms:7.3f} ms/render  {rays / ms / 1e3:8.1f} Mrays/s per rank  "
              f"launches/render {st['kernel_launches'] / n:.0f}  inner(ev0..ev1 of last render) {st['ms_total']:.3f} ms", flush=True)

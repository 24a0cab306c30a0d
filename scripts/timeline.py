"""Render rank 0 of WORLD on one GPU with per-launch events and print the lanes' traversal-launch timeline."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = "/tmp/trace_timeline.txt"
os.environ["TRACE_CUDA_TIMELINE"] = path
import torch
import trace_jl_b200 as T
from trace_jl_b200 import distributed as D
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 12
_s = torch.cuda.Stream(device=0)
torch.cuda.set_stream(_s)
ctx = T.Context(0, stream=_s.cuda_stream)
ctx.set_option("lanes", lanes)
scene, camera, spp, depth = bench.build_scene(T, "tess-1M")
H, W = camera.film.pixels.shape[:2]
film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
for i in range(3):
    D.render_whitted_sharded(ctx, scene, camera, spp, depth, 7 + i, film, 0, world, reduce=False)
ctx.set_option("time_kernels", 1)
if os.path.exists(path):
    os.remove(path)
D.render_whitted_sharded(ctx, scene, camera, spp, depth, 99, film, 0, world, reduce=False)
print("ms_total", ctx.stats()["ms_total"])
rows = [l.split() for l in open(path) if not l.startswith("#")]
by_lane = {}
for lane, kind, a, b in rows:
    by_lane.setdefault(int(lane), []).append((int(kind), float(a), float(b)))
for lane in sorted(by_lane):
    print(f"lane {lane:2d}: " + "  ".join(f"{'ESGHPRDUC'[k]}[{a:5.2f},{b:5.2f}]" for k, a, b in by_lane[lane]))

"""Traversal-launch timeline of one SPPM iteration (lane 0 = camera pass, lane 1 = photon pass)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = "/tmp/trace_timeline.txt"
os.environ["TRACE_CUDA_TIMELINE"] = path
import torch, trace_jl_b200 as T
from trace_jl_b200 import distributed as D
name = sys.argv[1] if len(sys.argv) > 1 else "caustic_moving"
scene, camera, kw = getattr(T.scenes, name)()
_s = torch.cuda.Stream(device=0)
torch.cuda.set_stream(_s)
ctx = T.Context(0, stream=_s.cuda_stream)
sess = D.SPPMSession(ctx, scene, camera, kw["initial_search_radius"], kw["max_depth"], kw.get("photons_per_iteration", -1))
for _ in range(3):
    sess.step()
ctx.synchronize()
ctx.set_option("time_kernels", 1)
if os.path.exists(path):
    os.remove(path)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sess.step()
e1.record()
torch.cuda.synchronize()
print(name, "iteration", e0.elapsed_time(e1), "ms")
sess.image()
for l in open(path):
    print(l.rstrip())

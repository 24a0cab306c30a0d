"""SPPM iterations/s on a named scene (GPU box): python scripts/sppm_bench.py shadows|caustic_glass|caustic_moving [iters] [res]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, trace_jl_b200 as T
from trace_jl_b200 import distributed as D
name = sys.argv[1] if len(sys.argv) > 1 else "shadows"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
res = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kw_res = dict(resolution=res) if res else {}
if os.environ.get("BUILDER"):
    kw_res.update(builder=os.environ["BUILDER"], max_node_primitives=int(os.environ.get("MNP", "1")))
scene, camera, kw = getattr(T.scenes, name)(**kw_res)
if os.environ.get("DEPTH"):
    kw["max_depth"] = int(os.environ["DEPTH"])
torch.cuda.set_device(0)
_s = torch.cuda.Stream(device=0)
torch.cuda.set_stream(_s)
ctx = T.Context(0, stream=_s.cuda_stream)
ctx.set_option("persist", int(os.environ.get("PERSIST", "0")))
ctx.set_option("sppm_lanes", int(os.environ.get("SPPM_LANES", "0")))
sess = D.SPPMSession(ctx, scene, camera, kw["initial_search_radius"], kw["max_depth"], kw.get("photons_per_iteration", -1))
for _ in range(2):
    sess.step()
torch.cuda.synchronize()
ctx.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    sess.step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
st = ctx.stats()
print(f"{name} depth={kw['max_depth']} builder={os.environ.get('BUILDER', 'reference')}/{os.environ.get('MNP', '1')} sppm_lanes={os.environ.get('SPPM_LANES', '0')} persist={os.environ.get('PERSIST', '0')}: {iters / ms * 1e3:.2f} it/s  ({ms / iters:.3f} ms/it)  rays/it extend {st['rays_extend'] / iters:.0f} shadow {st['rays_shadow'] / iters:.0f} "
      f"deposits/it {st['sppm_deposits'] / iters:.0f} launches/it {st['kernel_launches'] / iters:.1f} photons/it {sess.photons}")
if os.environ.get("COUNT_NODES"):
    ctx.set_option("count_nodes", 1)
    ctx.reset_stats()
    sess.step()
    ctx.synchronize()
    st = ctx.stats()
    nr = st["rays_extend"] + st["rays_shadow"]
    print(f"nodes/ray {st['nodes_visited'] / nr:.1f}  prims/ray {st['prims_tested'] / nr:.2f}  rays {nr}  prims_tested {st['prims_tested']}")
    ctx.set_option("count_nodes", 0)
img = sess.image()
print("image mean", float(img.mean()), "max", float(img.max()))
sess.close()

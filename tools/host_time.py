"""Development aid: host-side (CPU) time per section of one end-to-end C2 step through the public API; launches are
asynchronous, so this is what the Python / ctypes / autograd path costs when the GPU is not the limiter."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, F = (8, 5) if "--c3" in sys.argv else (1, 1)
o = bench.Ours(dev, 0, B, F, 2)
S, render, mods, host, cw = o.S, o.render, o.mods, o.host, o.cw
d = {k: host[0][k].to(dev) for k in bench.H2D_KEYS}
m = mods[0]
T = {}
def tick(name, t0):
    t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1

def step():
    t = time.perf_counter()
    sdf = d["sdf"].detach().requires_grad_(True); col = d["color"].detach().requires_grad_(True); sem = d["semantic"].detach().requires_grad_(True)
    t = tick("leaves", t)
    total, terms, _ = render(m, d["locs"], sdf, col, d["normal"], sem, d["view"], d["intr"], images_depth=d["t_depth"],
                             images_color=d["t_color"], target2d_label=d["t_label"], weight_semantic_class=cw, voxelsize=S.VOXELSIZE)
    t = tick("render", t)
    total.backward()
    t = tick("backward", t)

for _ in range(20): step()
torch.cuda.synchronize(); T.clear()
n = 200
for i in range(n):
    step()
    if i % 20 == 19: torch.cuda.synchronize()
print({k: "%.1f us" % (v / n * 1e6) for k, v in T.items()})
pr = cProfile.Profile(); pr.enable()
for i in range(n):
    step()
    if i % 20 == 19: torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

"""Small fixed workload for ncu: a few fwd+bwd passes of config C3 (or C2 with --c2) through the public module."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spsg_b200 import synthetic as S
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views

dev = torch.device("cuda", 0)
B, F = (1, 1) if "--c2" in sys.argv else (8, 5)
iters = 3
batch, t = scene_tensors(list(range(B)), dev)
n = t["locs"].shape[0]
_, _, view, intr = views(B, F, dev, seed=0)
mine = RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                   max_num_frames=F, max_num_locs_per_sample=(n + B - 1) // B + 1000, device=dev)
sdf = t["sdf"].clone().requires_grad_(True); sem = t["semantic"].clone().requires_grad_(True)
col = t["color"].clone().requires_grad_(True); nrm = t["normal"].clone().requires_grad_(True)
grads = None
for i in range(iters):
    out = mine(t["locs"], sdf, col, nrm, sem, view, intr)
    if grads is None:
        grads = [torch.randn_like(o) for o in out]
    torch.autograd.backward(out, grads)
torch.cuda.synchronize()
print("done", float(out[1][out[1] != -float('inf')].mean()))

"""Debug aid: read the forward's workspace maps back and compare with a numpy recomputation."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spsg_b200 import synthetic as S
from spsg_b200 import raycast_rgbd_cuda as rc
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views
dev = torch.device("cuda", 0)
batch, t = scene_tensors([0], dev)
n = t["locs"].shape[0]
_, _, view, intr = views(1, 1, dev, seed=0)
m = RaycastRGBD(1, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, max_num_locs_per_sample=n, device=dev)
m(t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view, intr)
torch.cuda.synchronize()
ws = rc._workspaces[("cuda", 0)].cpu().numpy()
Dz, Dy, Dx = S.DIMS_ZYX
al = lambda v: (v + 255) // 256 * 256
dense_b = al(Dz*Dy*Dx*4); skip_b = al((Dz//4)*(Dy//4)*(Dx//4)); wpr = (Dx+31)//32; vbit_b = al(Dz*Dy*wpr*8)
dense = ws[:Dz*Dy*Dx*4].view(np.float32).reshape(Dz, Dy, Dx)
skip = ws[dense_b:dense_b + (Dz//4)*(Dy//4)*(Dx//4)].reshape(Dz//4, Dy//4, Dx//4)
vb = ws[dense_b+skip_b: dense_b+skip_b+Dz*Dy*wpr*8].view(np.uint32).reshape(Dz, Dy, wpr, 2)
sdf, _ = S.sdf_volume(0)
P = np.abs(sdf) < 3
print("dense present match:", np.array_equal(~np.isnan(dense), P), "values match:", np.array_equal(dense[P], sdf[P]))
V = np.zeros_like(P); V[:-1,:-1,:-1] = P[:-1,:-1,:-1]&P[1:,:-1,:-1]&P[:-1,1:,:-1]&P[1:,1:,:-1]&P[:-1,:-1,1:]&P[1:,:-1,1:]&P[:-1,1:,1:]&P[1:,1:,1:]
bits = ((vb[..., None, :] >> np.arange(32, dtype=np.uint32)[None, None, None, :, None]) & 1).astype(bool)   # z,y,w,32,2
A = bits[..., 0].reshape(Dz, Dy, wpr*32)[:, :, :Dx]; B = bits[..., 1].reshape(Dz, Dy, wpr*32)[:, :, :Dx]
print("V match:", np.array_equal(A | B, V), "V count", V.sum(), "gpu", (A|B).sum(), "pos", (A&~B).sum(), "neg", (~A&B).sum(), "mixed", (A&B).sum())
def blocks(V, s): return V.reshape(Dz//s, s, Dy//s, s, Dx//s, s).any(axis=(1,3,5))
up = lambda a, f: a.repeat(f,0).repeat(f,1).repeat(f,2)
lvl = np.where(blocks(V,4), 0, np.where(up(blocks(V,8),2), 1, np.where(up(blocks(V,16),4), 2, np.where(up(blocks(V,32),8), 3, 4))))
print("level match:", np.array_equal(lvl, skip), "gpu hist", np.bincount(skip.ravel(), minlength=6), "ref hist", np.bincount(lvl.ravel(), minlength=6))

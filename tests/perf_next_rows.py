"""Device time of the producer / consumer ops next to the raycast (SURVEY 8f-1 and rows a11-a13 un-fused) against the
reference's literal PyTorch expressions (oracle/losses_ref.py), C3-sized inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/: it times the checkers next to the product)
sys.path.insert(0, ROOT)
import torch
from oracle import losses_ref as R
from spsg_b200 import synthetic as S, losses as L
from spsg_b200.normals import compute_normals_sparse
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views

dev = torch.device("cuda", 0)

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

B, F = 8, 1
_, t = scene_tensors(list(range(B)), dev)
n = t["locs"].shape[0]
_, _, view, intr = views(B, F, dev, seed=0)
tr = torch.inverse(view).contiguous()
w = torch.randn(n, 3, device=dev)
def fb(fn):
    def run():
        s = t["sdf"].clone().requires_grad_(True)
        (fn(s) * w).sum().backward()
    return run
print("normals fwd+bwd, %d voxels in %d chunks: reference expression %.0f us, fused op %.0f us" % (
    n, B, timeit(fb(lambda s: R.compute_normals_sparse(t["locs"], s, S.DIMS_ZYX, tr))),
    timeit(fb(lambda s: compute_normals_sparse(t["locs"], s, S.DIMS_ZYX, tr, num_chunks=B)))))
rc = RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                 max_num_locs_per_sample=n // B + 1000, device=dev)
with torch.no_grad():
    color, depth, normal, sem = [x.clone() for x in rc(t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view, intr)]
g = torch.Generator(device=dev).manual_seed(1)
tdepth = torch.rand(B, 1, S.HEIGHT, S.WIDTH, device=dev, generator=g) + 0.5
tcolor = torch.rand(B, S.HEIGHT, S.WIDTH, 3, device=dev, generator=g)
label = torch.randint(0, 15, (B, S.HEIGHT, S.WIDTH, 1), device=dev, generator=g).to(torch.uint8)
cw = torch.tensor(S.CLASS_WEIGHTS, device=dev)
def three(fd, fc, fs):
    def run():
        c = color.clone().requires_grad_(True); d = depth.clone().requires_grad_(True); s = sem.clone().requires_grad_(True)
        (fd(d) + fc(c) + fs(s)).backward()
    return run
print("depth L1 + colour L1 + 2D CE fwd+bwd on %d rendered images: reference expressions %.0f us, stand-alone ops %.0f us" % (
    B, timeit(three(lambda d: R.depth_l1_loss(d, tdepth, S.VOXELSIZE), lambda c: R.compute_2dcolor_loss(c, tcolor, None), lambda s: R.semantic_2d_ce_loss(s, label, cw))),
    timeit(three(lambda d: L.depth_l1_loss(d, tdepth, S.VOXELSIZE), lambda c: L.color_l1_loss(c, tcolor), lambda s: L.semantic_2d_ce_loss(s, label, cw)))))
def one():
    c = color.clone().requires_grad_(True); d = depth.clone().requires_grad_(True); s = sem.clone().requires_grad_(True)
    L.losses_2d(c, d, s, images_depth=tdepth, images_color=tcolor, target2d_label=label, weight_semantic_class=cw, voxelsize=S.VOXELSIZE)[0].backward()
print("  same three terms in one pass (losses_2d): %.0f us" % timeit(one))

# ---- Depth2Normals: reference extension pipeline (host check per fill round) vs one enqueued native call
from oracle import ref_driver
from spsg_b200.depth_utils import Depth2Normals
if ref_driver.depth_available():
    Bd, Hd, Wd = 8, S.HEIGHT, S.WIDTH
    g2 = torch.Generator().manual_seed(0)
    base = 1.5 + 0.3 * torch.rand(Bd, 1, Hd, Wd, generator=g2)
    base[torch.rand(Bd, 1, Hd, Wd, generator=g2) < 0.03] = 0.0
    base = base.to(dev)
    intr8 = torch.tensor([list(S.INTRINSICS)] * Bd, device=dev)
    mod = Depth2Normals(Bd, Wd, Hd, S.DEPTH_MIN, S.DEPTH_MAX, device=dev)
    filt, cam, nrm = torch.zeros_like(base), torch.zeros(Bd, Hd, Wd, 3, device=dev), torch.zeros(Bd, Hd, Wd, 3, device=dev)
    print("Depth2Normals, %d frames %dx%d with 3%% holes: reference extension %.0f us, fused pipeline %.0f us" % (
        Bd, Wd, Hd, timeit(lambda: ref_driver.ref_depth2normals(base.clone(), intr8, filt, cam, nrm)),
        timeit(lambda: mod(base.clone(), intr8))))

# ---- producer glue (train.py:494-509): dense heads -> locs + payloads, forward + backward
from spsg_b200 import sparsify
dz, dy, dx = S.DIMS_ZYX
l = t["locs"]
dense_sdf = torch.full((B, 1, dz, dy, dx), 10.0, device=dev)
dense_sdf[l[:, 3], 0, l[:, 0], l[:, 1], l[:, 2]] = t["sdf"][:, 0]
dense_col = torch.rand(B, 3, dz, dy, dx, device=dev)
dense_sem = torch.randn(B, 14, dz, dy, dx, device=dev)
def literal():
    heads = [h.clone().requires_grad_(True) for h in (dense_sdf, dense_col, dense_sem)]
    locs = torch.nonzero(torch.abs(heads[0].detach()[:, 0]) < S.TRUNCATION)
    locs = torch.cat([locs[:, 1:], locs[:, :1]], 1)
    vals = [h[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]] for h in heads]
    sum(v.sum() for v in vals).backward()
def fused():
    heads = [h.clone().requires_grad_(True) for h in (dense_sdf, dense_col, dense_sem)]
    out = sparsify.sparsify_predictions(heads[0], S.TRUNCATION, None, heads[1], heads[2])
    sum(v.sum() for v in out[1:]).backward()
def clones():
    [h.clone().requires_grad_(True) for h in (dense_sdf, dense_col, dense_sem)]
base = timeit(clones)
print("dense heads -> locs + sdf/colour/semantic values, fwd+bwd, %d chunks (%d voxels): reference expressions %.0f us, "
      "sparsify ops %.0f us (both minus %.0f us of head clones)" % (B, n, timeit(literal) - base, timeit(fused) - base, base))

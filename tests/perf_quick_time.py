"""Quick device timing of mine vs the reference extension on configs C2 / C3 (development aid, not the bench)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/: it times the checkers next to the product)
sys.path.insert(0, ROOT)
from spsg_b200 import synthetic as S
from spsg_b200.raycast_rgbd import RaycastRGBD
from oracle import ref_driver as refdriver
from tests.common import scene_tensors, views

dev = torch.device("cuda", 0)

def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us

def run(B, F, flags_list=(0, 1, 2, 3)):
    batch, t = scene_tensors(list(range(B)), dev)
    n = t["locs"].shape[0]
    _, _, view, intr = views(B, F, dev, seed=0)
    rays = B * F * S.WIDTH * S.HEIGHT
    print("== B=%d F=%d N=%d rays=%d" % (B, F, n, rays))
    mine = RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                       max_num_frames=F, max_num_locs_per_sample=(n + B - 1) // B + 1000, device=dev)
    sdf = t["sdf"].clone().requires_grad_(True); sem = t["semantic"].clone().requires_grad_(True)
    col = t["color"].clone().requires_grad_(True); nrm = t["normal"].clone().requires_grad_(True)
    for fl in flags_list:
        mine.flags = fl
        def fwd():
            with torch.no_grad():
                mine(t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view, intr)
        us = timeit(fwd)
        print("mine fwd flags=%d: %.1f us  %.2f Grays/s" % (fl, us, rays / us / 1e3))
    mine.flags = 0
    out = mine(t["locs"], sdf, col, nrm, sem, view, intr)
    print("hit rate %.3f, max pixels/voxel %d" % ((out[1] != -float('inf')).float().mean().item(), mine.mapping3dto2d_num.max().item()))
    grads = [torch.randn_like(o) for o in out]
    def fb():
        o = mine(t["locs"], sdf, col, nrm, sem, view, intr)
        torch.autograd.backward(o, grads)
    us = timeit(fb)
    print("mine fwd+bwd (autograd): %.1f us  %.2f Grays/s" % (us, rays / us / 1e3))
    from spsg_b200 import _native as N
    N.timing_read(0); N.timing_read(1); N.timing_enable(True)
    for _ in range(20): fb()
    torch.cuda.synchronize(); N.timing_enable(False)
    f_ms, f_n = N.timing_read(0); g_ms, g_n = N.timing_read(1)
    print("kernel device time: raycast_forward %.1f us, backward_gather %.1f us" % (f_ms / f_n * 1e3, g_ms / g_n * 1e3))
    from spsg_b200 import raycast_rgbd_cuda as rc
    dims = [B, 64, 64, 128, n]
    def bwd():
        rc.backward(grads[0], grads[1], grads[2], grads[3], mine.sparse_mapping, mine.mapping3dto2d, mine.mapping3dto2d_num, dims,
                    mine.d_color, mine.d_depth, mine.d_normal, mine.d_semantic, views_per_chunk=F)
    us = timeit(bwd)
    print("mine bwd only: %.1f us" % us)
    if refdriver.available():
        nmax = (n + B - 1) // B + 1000
        ref = refdriver.RefRaycaster(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, nmax, 64, device=dev)
        def rf():
            for f in range(F):
                sel = torch.arange(B, device=dev) * F + f
                ref.forward(t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view[sel].contiguous(), intr[sel].contiguous())
        us = timeit(rf, iters=10, warm=2)
        print("ref fwd (x%d views, max_locs=%d): %.1f us  %.3f Grays/s" % (F, nmax, us, rays / us / 1e3))
        g1 = [g[:B].contiguous() for g in grads]
        def rb():
            for f in range(F):
                ref.backward(*g1)
        us = timeit(rb, iters=10, warm=2)
        print("ref bwd (x%d): %.1f us" % (F, us))

run(1, 1)
run(8, 5, flags_list=(0, 3))

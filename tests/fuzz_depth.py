"""Randomised check of the depth-frame utilities against the compiled reference extension (not collected by pytest).
Random frame sizes, hole densities and depth ranges; holes stay sparse enough that every 11x11 window keeps at least two
valid depths (below that the reference reads out of bounds).  usage: python tests/fuzz_depth.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import ref_driver as refdriver
from spsg_b200.depth_utils import Depth2Normals

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
assert refdriver.depth_available(), "oracle/_ref depth extension not built"
bits = lambda t: t.contiguous().view(torch.int32)
bad = 0
t0 = time.time()
for c in range(cases):
    rng = np.random.default_rng(seed0 * 104729 + c)
    g = torch.Generator().manual_seed(int(rng.integers(0, 1 << 30)))
    b, h, w = int(rng.integers(1, 4)), int(rng.integers(12, 120)), int(rng.integers(12, 160))
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    depth = torch.stack([float(rng.uniform(0.5, 4)) + 0.01 * float(rng.uniform(-1, 1)) * xx + 0.01 * float(rng.uniform(-1, 1)) * yy +
                         0.2 * torch.sin(xx / float(rng.uniform(3, 20))) * torch.cos(yy / float(rng.uniform(3, 20))) +
                         0.02 * torch.randn(h, w, generator=g) for _ in range(b)])
    depth = depth.clamp_min(0.05)
    depth[torch.rand(b, h, w, generator=g) < float(rng.uniform(0, 0.3))] = 0.0
    if rng.random() < 0.5:
        y0, x0 = int(rng.integers(0, h - 6)), int(rng.integers(0, w - 6))
        depth[:, y0:y0 + int(rng.integers(1, 6)), x0:x0 + int(rng.integers(1, 6))] = 0.0
    valid = torch.nn.functional.avg_pool2d((depth != 0).float()[:, None], 11, stride=1, padding=5, divisor_override=1)
    if float(valid.min()) < 1.5:
        continue
    depth = depth[:, None].contiguous().to(dev)
    intr = torch.tensor([[float(rng.uniform(50, 300)), float(rng.uniform(50, 300)), w / 2 - 0.5, h / 2 - 0.5]] * b, device=dev)
    iters = int(rng.choice([0, 2, 4, 40]))
    d_mine, d_ref = depth.clone(), depth.clone()
    mod = Depth2Normals(b, w, h, 0.1 / 0.02, 6.0 / 0.02, max_num_fill_iters=iters, device=dev)
    got = mod(d_mine, intr)
    filt, cam, nrm = torch.zeros_like(depth), torch.zeros(b, h, w, 3, device=dev), torch.zeros(b, h, w, 3, device=dev)
    want = refdriver.ref_depth2normals(d_ref, intr, filt, cam, nrm, max_num_fill_iters=iters)
    ok = (got is None) == (want is None) and torch.equal(bits(d_mine), bits(d_ref))
    if ok and got is not None:
        ok = torch.equal(bits(got), bits(want)) and torch.equal(bits(mod.camspace), bits(cam)) and torch.equal(bits(mod.filter_helper), bits(filt))
    if not ok:
        bad += 1
        print("MISMATCH case %d: b %d h %d w %d iters %d none %s/%s" % (c, b, h, w, iters, got is None, want is None), flush=True)
print("%d cases, %d mismatching, %.1f s" % (cases, bad, time.time() - t0))
sys.exit(1 if bad else 0)

"""Randomised check of the fused raycast + 2D losses (run as a script for long hunts; tests/test_gpu_fuzz_slice.py runs a seeded slice under pytest): random soups, cameras, views per chunk,
targets and term switches; loss values and voxel gradients against the literal expressions applied to the un-fused rendering
of the same module.  usage: python tests/fuzz_fused.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import losses_ref as R
from tests.test_gpu_adversarial import _batch, _cameras
from spsg_b200.losses import render_with_2d_losses
from spsg_b200.raycast_rgbd import RaycastRGBD

def run(cases=100, seed0=0, dev=None):
    """Returns the number of mismatching cases."""
    dev = dev or torch.device("cuda", 0)
    bad = skipped = 0
    t0 = time.time()
    for c in range(cases):
        rng = np.random.default_rng(seed0 * 15485863 + c)
        g = torch.Generator().manual_seed(int(rng.integers(0, 1 << 30)))
        dims = tuple(int(v) for v in rng.integers(6, 40, 3))
        B, F = int(rng.integers(1, 4)), int(rng.integers(1, 3))
        kinds = ["noise", "blocky", "special"]
        specs = [(int(rng.integers(0, 1 << 30)), kinds[int(rng.integers(0, 2))]) if rng.random() > 0.15 else None for _ in range(B)]
        if all(s is None for s in specs):
            specs[0] = (3, "blocky")
        w, h = int(rng.integers(8, 70)), int(rng.integers(8, 50))
        t, n = _batch(dims, specs, dev)
        I = B * F
        view = _cameras(dims, I, int(rng.integers(0, 1 << 30)), dev)
        intr = torch.tensor([[float(w), float(w), (w - 1) / 2, (h - 1) / 2]] * I, device=dev)
        rc = RaycastRGBD(B, dims, w, h, 0.0, 150.0, 50.0, float(rng.choice([0.9, 0.5, 1.3])), max_num_frames=F,
                         max_num_locs_per_sample=n, device=dev)
        use_d, use_c, use_s = (bool(rng.random() < 0.8) for _ in range(3))
        if not (use_d or use_c or use_s):
            use_d = True
        images_depth = torch.rand(I, 1, h, w, generator=g) * 2.0
        images_depth[torch.rand(I, 1, h, w, generator=g) < 0.2] = 0.0
        images_color = torch.rand(I, h, w, 3, generator=g)
        label = torch.randint(0, 15, (I, h, w, 1), generator=g).to(torch.uint8)
        wc = (torch.rand(I, 1, h, w, generator=g) + 0.5) if rng.random() < 0.5 else None
        cw = (torch.rand(14, generator=g) + 0.1) if rng.random() < 0.7 else None
        images_depth, images_color, label = images_depth.to(dev), images_color.to(dev), label.to(dev)
        wc = None if wc is None else wc.to(dev)
        cw = None if cw is None else cw.to(dev)
        wts = [float(x) for x in rng.uniform(0.2, 2.0, 3)]
        leafs = lambda: [t[k].clone().requires_grad_(True) for k in ("sdf", "color", "semantic")]
        sdf, col, sem = leafs()
        r_color, r_depth, _, r_sem = rc(t["locs"], sdf, col, t["normal"], sem, view, intr)
        if int(rc.mapping3dto2d_num[:n * F].max()) > 64:
            skipped += 1
            continue
        total = 0.0
        if use_d: total = total + wts[0] * R.depth_l1_loss(r_depth, images_depth, 0.02)
        if use_c: total = total + wts[1] * R.compute_2dcolor_loss(r_color, images_color, wc)
        if use_s: total = total + wts[2] * R.semantic_2d_ce_loss(r_sem, label, cw)
        if not torch.isfinite(total):
            skipped += 1   # a term over an empty pixel set is NaN in both
            continue
        total.backward()
        want = [x.grad if x.grad is not None else torch.zeros_like(x) for x in (sdf, col, sem)]
        sdf2, col2, sem2 = leafs()
        total2, _, _ = render_with_2d_losses(rc, t["locs"], sdf2, col2, t["normal"], sem2, view, intr,
                                             images_depth=images_depth if use_d else None, images_color=images_color if use_c else None,
                                             weight_color=wc if use_c else None, target2d_label=label if use_s else None,
                                             weight_semantic_class=cw, voxelsize=0.02, weight_depth_loss=wts[0],
                                             weight_color_loss=wts[1], weight_semantic_loss=wts[2])
        total2.backward()
        ok = abs(float(total2.detach()) - float(total.detach())) <= 1e-5 * max(1.0, abs(float(total.detach())))
        for a, b in zip((sdf2.grad, col2.grad, sem2.grad), want):
            a = a if a is not None else torch.zeros_like(b)
            ok = ok and float((a - b).abs().max()) <= 1e-3 * (float(b.abs().max()) + 1e-12) + 1e-9
        if not ok:
            bad += 1
            print("MISMATCH case %d: dims %s B %d F %d img %dx%d terms %s total %g vs %g" % (
                c, dims, B, F, w, h, (use_d, use_c, use_s), float(total2.detach()), float(total.detach())), flush=True)
    print("%d cases, %d skipped, %d mismatching, %.1f s" % (cases, skipped, bad, time.time() - t0))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)

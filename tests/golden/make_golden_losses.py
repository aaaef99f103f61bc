#!/usr/bin/env python
"""Generate tests/golden/losses_ref.npz by running the reference's OWN Python code (CPU is enough).

    python tests/golden/make_golden_losses.py        (needs /root/reference; run in the build container)

`/root/reference/torch/loss.py` is imported unmodified; its module-level `import data_util` pulls in image/PLY I/O
packages that are not installed here and that none of the functions used below touch, so an empty stand-in module is
registered for it first.  Recorded: `loss.compute_normals_sparse` (loss.py:285-306) with the gradient of a fixed linear
functional w.r.t. the SDF values, and `loss.compute_2dcolor_loss` (loss.py:246-257) with and without per-pixel weights and
its gradient.  They pin oracle/losses_ref.py (tests/test_oracle_cpu.py) and the CUDA ops (tests/test_gpu_normals.py,
tests/test_gpu_losses.py) where /root/reference does not exist."""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SPSG_REFERENCE_TORCH", "/root/reference/torch")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "losses_ref.npz")


def main():
    sys.modules.setdefault("data_util", types.ModuleType("data_util"))
    sys.path.insert(0, REF)
    import loss as ref_loss

    g = torch.Generator().manual_seed(20261018)
    # ---- normals: two small chunks, band voxels of a tilted plane + a blob, camera-like rotations
    dims = (12, 10, 14)
    z, y, x = torch.meshgrid(torch.arange(dims[0]), torch.arange(dims[1]), torch.arange(dims[2]), indexing="ij")
    locs, vals = [], []
    for b in range(2):
        d = (0.6 * z + 0.3 * y - 0.2 * x - 3.7 - b).float()
        d = torch.minimum(d, ((x - 8.2) ** 2 + (y - 4.6) ** 2 + (z - 5.1 - b) ** 2).float().sqrt() - 2.8)
        d = d.clamp(-3, 3) + 0.05 * torch.randn(dims, generator=g)
        m = d.abs() < 2.5
        l = torch.nonzero(m)
        locs.append(torch.cat([l, torch.full((l.shape[0], 1), b, dtype=torch.long)], 1))
        vals.append(d[m].reshape(-1, 1))
    locs, vals = torch.cat(locs), torch.cat(vals)
    q = torch.linalg.qr(torch.randn(2, 3, 3, generator=g))[0]
    transform = torch.eye(4).repeat(2, 1, 1)
    transform[:, :3, :3] = q
    transform[:, :3, 3] = torch.randn(2, 3, generator=g)
    w = torch.randn(locs.shape[0], 3, generator=g)
    out = {}
    for name, tr in (("normals_t", transform), ("normals_id", None)):
        v = vals.clone().requires_grad_(True)
        # the reference moves the zero volume to sdf_vals.device and indexes with .cuda()-free code: CPU works as is
        n = ref_loss.compute_normals_sparse(locs, v, dims, transform=tr)
        (n * w).sum().backward()
        out[name] = n.detach().numpy()
        out[name + "_dsdf"] = v.grad.numpy()
    out.update(n_locs=locs.numpy(), n_sdf=vals.numpy(), n_dims=np.array(dims), n_transform=transform.numpy(), n_w=w.numpy())

    # ---- colour L1: rendering with -inf holes, target, optional per-pixel weight (B,1,H,W)
    B, H, W = 2, 12, 16
    pred = torch.rand(B, H, W, 3, generator=g)
    hole = torch.rand(B, H, W, generator=g) < 0.3
    pred[hole] = -float("inf")
    tgt = torch.rand(B, H, W, 3, generator=g)
    wt = torch.rand(B, 1, H, W, generator=g) * 2
    for name, ww in (("color_w", wt), ("color_now", None)):
        p = pred.clone().requires_grad_(True)
        l = ref_loss.compute_2dcolor_loss(p, tgt, ww)
        l.backward()
        out[name] = np.float32(l.item())
        gr = p.grad.clone()
        gr[hole] = 0          # gradient of the holes is NaN/0 garbage of inf arithmetic in the masked path: not compared
        out[name + "_grad"] = gr.numpy()
    out.update(c_pred=pred.numpy(), c_tgt=tgt.numpy(), c_weight=wt.numpy())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", locs.shape[0], "voxels")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference extension (oracle/_ref) on a CUDA box.

    python tests/golden/make_golden.py [outdir]        (run on the GPU box via gpurun; copy *.npz into tests/golden/)

Each file holds a tiny seeded scene (inputs) and what the reference's native forward / backward produced for it.
They pin the CPU restatement (tests/test_oracle_cpu.py::test_golden_vectors_from_reference_extension) and the CUDA
path (tests/test_gpu_parity.py::test_golden_vectors_bit_exact)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver  # noqa: E402
from spsg_b200 import synthetic as S  # noqa: E402


def scene(seed, dims=(32, 32, 32)):
    rng = np.random.default_rng(seed)
    dz, dy, dx = dims
    z, y, x = np.meshgrid(np.arange(dz), np.arange(dy), np.arange(dx), indexing="ij")
    c = np.array([15.3, 16.1, 17.2]) + rng.uniform(-1, 1, 3)
    d_sphere = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - 8.6
    d_floor = z - 5.4
    sdf = np.clip(np.minimum(d_sphere, d_floor), -3, 3).astype(np.float32)
    mask = np.abs(sdf) < 3
    locs = np.argwhere(mask).astype(np.int64)
    locs = np.ascontiguousarray(np.concatenate([locs, np.zeros((locs.shape[0], 1), np.int64)], 1))
    n = locs.shape[0]
    color = (rng.integers(0, 256, (n, 3)) / 255.0).astype(np.float32)
    normal = rng.standard_normal((n, 3)).astype(np.float32)
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    normal[::7] = 0.0
    semantic = rng.integers(-8, 9, (n, 14)).astype(np.float32)
    return dims, locs, sdf[mask].reshape(n, 1), color, normal.astype(np.float32), semantic


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    dev = torch.device("cuda", 0)
    w, h = 48, 40
    for name, seed, inc, thresh, dmin in (("sphere_a", 1, 0.9, 45.45, 5.0), ("sphere_b", 2, 0.45, 2.0, 0.5)):
        dims, locs, sdf, color, normal, semantic = scene(seed)
        view = S.look_at((16.0 + 3 * seed, -24.0, 30.0), (16.0, 16.0, 12.0))[None].astype(np.float32)
        intr = np.array([[60.0, 60.5, 23.5, 19.5]], np.float32)
        t = [torch.from_numpy(a).to(dev) for a in (locs, sdf, color, normal, semantic, view, intr)]
        ref = ref_driver.RefRaycaster(1, dims, w, h, dmin, 100.0, thresh, inc, locs.shape[0], 64, device=dev)
        img = [o.clone() for o in ref.forward(*t)]
        g = torch.Generator(device="cpu").manual_seed(seed)
        grads = [torch.randint(-4, 5, o.shape, generator=g).float().to(dev) for o in img]
        d = [x.clone() for x in ref.backward(*grads)]
        n = locs.shape[0]
        assert int(ref.mapping3dto2d_num[:n].max()) <= 64
        np.savez_compressed(
            os.path.join(out, name + ".npz"), dims_zyx=np.array(dims), locs=locs, sdf=sdf, color=color, normal=normal,
            semantic=semantic, view=view, intr=intr, depth_min=dmin, depth_max=100.0, thresh=thresh, inc=inc,
            num_chunks=1, color_img=img[0].cpu().numpy(), depth=img[1].cpu().numpy(), normal_img=img[2].cpu().numpy(),
            semantic_img=img[3].cpu().numpy(), num=ref.mapping3dto2d_num[:n].cpu().numpy(),
            g_color=grads[0].cpu().numpy(), g_depth=grads[1].cpu().numpy(), g_normal=grads[2].cpu().numpy(),
            g_semantic=grads[3].cpu().numpy(), d_color=d[0].cpu().numpy(), d_depth=d[1].cpu().numpy(),
            d_normal=d[2].cpu().numpy(), d_semantic=d[3].cpu().numpy())
        hit = torch.isfinite(img[1]).float().mean().item()
        print(name, "voxels", n, "hit rate %.3f" % hit, "max px/voxel", int(ref.mapping3dto2d_num[:n].max()))


if __name__ == "__main__":
    main()

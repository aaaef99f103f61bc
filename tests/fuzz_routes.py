"""Randomised check of the two alternative routes into the forward against the plain one (int64 rows -> fill -> index):
(a) rows packed on the host into uint32 cell indices (spsg_pack_locs_host + SPSG_FLAG_PACKED_LOCS),
(b) voxel index + SDF brick written by the compaction pass (spsg_sparsify_locs_indexed + SPSG_FLAG_INDEX_PREBUILT).
Renderings, voxel index and voxel -> pixel counters must be equal bit for bit.  usage: python tests/fuzz_routes.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from tests.test_gpu_adversarial import _cameras
from tests.common import count_bit_mismatch
from spsg_b200 import raycast_rgbd_cuda as rc, sparsify
from spsg_b200.raycast_rgbd import RaycastRGBD


def run(cases=100, seed0=0, dev=None):
    dev = dev or torch.device("cuda", 0)
    bad = compared = hits = 0
    t0 = time.time()
    for c in range(cases):
        rng = np.random.default_rng(seed0 * 100019 + c)
        g = torch.Generator().manual_seed(int(rng.integers(0, 1 << 30)))
        dims = tuple(int(v) for v in rng.integers(3, 60, 3))
        B, F = int(rng.integers(1, 4)), int(rng.integers(1, 3))
        trunc = float(rng.choice([3.0, 1.5, 2.5]))
        # a dense SDF head: smooth field + noise, so that |sdf| < truncation selects shells with zero crossings
        zz, yy, xx = torch.meshgrid(*[torch.arange(d, dtype=torch.float32) for d in dims], indexing="ij")
        heads = []
        for b in range(B):
            ctr = [float(rng.uniform(0.2, 0.8)) * d for d in dims]
            r = float(rng.uniform(0.15, 0.45)) * min(dims)
            heads.append(torch.sqrt((zz - ctr[0]) ** 2 + (yy - ctr[1]) ** 2 + (xx - ctr[2]) ** 2) - r + 0.3 * torch.randn(dims, generator=g))
        sdf = torch.stack(heads)[:, None].contiguous().to(dev)
        empty = (torch.rand(B, 1, *dims, generator=g) < 0.1).to(dev) if rng.random() < 0.5 else None
        col = torch.rand(B, 3, *dims, generator=g).to(dev)
        sem = torch.randn(B, 14, *dims, generator=g).to(dev)
        w, h = int(rng.integers(8, 90)), int(rng.integers(8, 70))
        inc = float(rng.choice([0.9, 0.5, 1.3, float(rng.uniform(0.2, 2.0))]))
        view = _cameras(dims, B * F, int(rng.integers(0, 1 << 30)), dev)
        f = float(rng.uniform(0.5, 1.4)) * w
        intr = torch.tensor([[f, f, (w - 1) / 2, (h - 1) / 2]] * (B * F), device=dev)
        n_max = int(np.prod(dims))
        mk = lambda: RaycastRGBD(B, dims, w, h, 0.0, 200.0, 50.0, inc, max_num_frames=F, max_num_locs_per_sample=n_max, device=dev)
        outs = []
        with torch.no_grad():
            # plain
            m = mk()
            locs, v_sdf, v_col, v_sem = sparsify.sparsify_predictions(sdf, trunc, empty, col, sem)
            if locs.shape[0] == 0:
                continue
            nrm = torch.nn.functional.normalize(torch.randn(locs.shape[0], 3, generator=g), dim=1).to(dev)
            o = m(locs, v_sdf, v_col, nrm, v_sem, view, intr)
            outs.append((o, m.sparse_mapping.clone(), m.mapping3dto2d_num[:locs.shape[0] * F].clone()))
            # packed on the host
            m = mk()
            cells = rc.pack_locs_host(locs.cpu(), B, dims, threads=int(rng.integers(1, 5))).to(dev)
            o = m(cells, v_sdf, v_col, nrm, v_sem, view, intr)
            outs.append((o, m.sparse_mapping.clone(), m.mapping3dto2d_num[:locs.shape[0] * F].clone()))
            # index + brick written by the compaction
            m = mk()
            locs2, v_sdf2, v_col2, v_sem2 = sparsify.sparsify_predictions(sdf, trunc, empty, col, sem, raycaster=m)
            took = m.workspace._prebuilt is not None
            o = m(locs2, v_sdf2, v_col2, nrm, v_sem2, view, intr)
            outs.append((o, m.sparse_mapping.clone(), m.mapping3dto2d_num[:locs.shape[0] * F].clone()))
        compared += 1
        hits += int((outs[0][0][1] != -float("inf")).sum())
        for name, alt in (("packed", outs[1]), ("prebuilt", outs[2])):
            mism = [count_bit_mismatch(a, b) for a, b in zip(outs[0][0], alt[0])]
            ok = not any(mism) and torch.equal(outs[0][1], alt[1]) and torch.equal(outs[0][2], alt[2])
            if not ok or not took:
                bad += 1
                print("MISMATCH case %d (%s): dims %s B %d F %d img %dx%d inc %g n %d -> %s index %s counters %s prebuilt-mark %s"
                      % (c, name, dims, B, F, w, h, inc, locs.shape[0], mism, torch.equal(outs[0][1], alt[1]),
                         torch.equal(outs[0][2], alt[2]), took), flush=True)
    print("%d cases, %d compared (%d hit pixels in all), %d mismatching, %.1f s" % (cases, compared, hits, bad, time.time() - t0))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)

"""Randomised checks of the next-row ops against the reference's literal PyTorch expressions (not collected by pytest):
sparsify (train.py:494-509), label maps (train.py:614-616), stand-alone 2D losses, compute_normals_sparse.
usage: python tests/fuzz_next_rows.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import losses_ref as R
from spsg_b200 import sparsify, losses as L, _native as N
from spsg_b200.normals import compute_normals_sparse

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
bad = 0
t0 = time.time()
def fail(c, what):
    global bad
    bad += 1
    print("MISMATCH case %d: %s" % (c, what), flush=True)
for c in range(cases):
    rng = np.random.default_rng(seed0 * 7919 + c)
    g = torch.Generator().manual_seed(int(rng.integers(0, 1 << 30)))
    # ---- sparsify
    B = int(rng.integers(1, 5)); dz, dy, dx = (int(v) for v in rng.integers(1, 40, 3))
    trunc = float(rng.choice([3.0, 1.0, 0.5, 2.5]))
    sdf = (torch.randn(B, 1, dz, dy, dx, generator=g) * 2.0).to(dev)
    sdf.view(-1)[::53] = float("nan"); sdf.view(-1)[7::61] = trunc
    empty = (torch.rand(B, 1, dz, dy, dx, generator=g) < 0.3).to(dev) if rng.random() < 0.5 else None
    heads = [(torch.randn(B, int(ch), dz, dy, dx, generator=g)).to(dev).requires_grad_(True) for ch in rng.integers(1, 20, int(rng.integers(1, 6)))]
    mask = torch.abs(sdf[:, 0]) < trunc
    if empty is not None:
        mask = mask & ~empty[:, 0]
    locs_ref = torch.nonzero(mask); locs_ref = torch.cat([locs_ref[:, 1:], locs_ref[:, :1]], 1)
    locs = sparsify.sparse_locs(sdf, trunc, empty)
    if not torch.equal(locs, locs_ref):
        fail(c, "sparse_locs %s" % ((B, dz, dy, dx),)); continue
    # the indexed variant (spsg_sparsify_locs_indexed): same rows, plus voxel index and SDF brick for every cell
    counted = sparsify.count_locs(sdf, trunc, empty)
    cells, n = sdf.numel(), counted.n
    index = torch.full((cells,), 12345, dtype=torch.int32, device=dev)
    brick = torch.zeros(cells, device=dev)
    locs_i = torch.empty(n, 4, dtype=torch.int64, device=dev)
    N.check(N.lib.spsg_sparsify_locs_indexed(N.ptr(counted.sdf), N.ptr(counted.empty), B, dz, dy, dx, trunc, N.ptr(counted.scratch),
                                             N.ptr(locs_i), n, N.ptr(index), N.ptr(brick), torch.cuda.current_stream().cuda_stream))
    want_index = torch.full((B, dz, dy, dx), -1, dtype=torch.int32, device=dev)
    want_index[locs_ref[:, 3], locs_ref[:, 0], locs_ref[:, 1], locs_ref[:, 2]] = torch.arange(n, dtype=torch.int32, device=dev)
    want_brick = torch.full((B, dz, dy, dx), -1, dtype=torch.int32, device=dev)  # 0xffffffff = absent
    want_brick[mask] = sdf[:, 0][mask].view(torch.int32)
    if not (torch.equal(locs_i, locs_ref) and torch.equal(index, want_index.view(-1)) and
            torch.equal(brick.view(torch.int32), want_brick.view(-1))):
        fail(c, "sparsify_locs_indexed %s" % ((B, dz, dy, dx),)); continue
    if locs.shape[0]:
        vals = sparsify.gather_dense(locs, *heads)
        vals = vals if isinstance(vals, tuple) else (vals,)
        ws = [torch.randn(v.shape, generator=g).to(dev) for v in vals]
        sum((v * w).sum() for v, w in zip(vals, ws)).backward()
        got = [h.grad.clone() for h in heads]
        for h in heads: h.grad = None
        ref_vals = [h[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]] for h in heads]
        sum((v * w).sum() for v, w in zip(ref_vals, ws)).backward()
        if not all(torch.equal(a, b) for a, b in zip(vals, ref_vals)) or not all(torch.equal(a, h.grad) for a, h in zip(got, heads)):
            fail(c, "gather_dense"); continue
    # ---- label maps + stand-alone losses
    I, h, w = int(rng.integers(1, 4)), int(rng.integers(1, 50)), int(rng.integers(1, 60))
    sem = (torch.randn(I, h, w, 14, generator=g) * 2).to(dev)
    miss = (torch.rand(I, h, w, generator=g) < 0.3).to(dev)
    sem[miss] = -float("inf")
    sem.view(-1, 14)[::11] = torch.nn.functional.one_hot(torch.arange(sem.view(-1, 14)[::11].shape[0]) % 14, 14).float().to(dev)
    cat = torch.cat((sem, torch.ones(sem.shape[:-1] + (1,), device=dev)), dim=-1)
    want = torch.max(cat, dim=-1)[1].to(torch.uint8)
    lab, hist = L.labels_from_render(sem, histogram=True)
    if not torch.equal(lab, want) or not torch.equal(hist, torch.bincount(want.reshape(-1).long(), minlength=15)):
        fail(c, "labels"); continue
    depth = (torch.rand(I, h, w, generator=g) * 100).to(dev); depth[miss] = -float("inf")
    color = torch.rand(I, h, w, 3, generator=g).to(dev); color[miss] = -float("inf")
    t_depth = (torch.rand(I, 1, h, w, generator=g) * 2).to(dev); t_depth[torch.rand(I, 1, h, w, generator=g).to(dev) < 0.2] = 0.0
    t_color = torch.rand(I, h, w, 3, generator=g).to(dev)
    label = torch.randint(0, 15, (I, h, w, 1), generator=g).to(torch.uint8).to(dev)
    cw = (torch.rand(14, generator=g) + 0.1).to(dev)
    leafs = [x.clone().requires_grad_(True) for x in (color, depth, sem)]
    leafs2 = [x.clone().requires_grad_(True) for x in (color, depth, sem)]
    tot_ref = R.depth_l1_loss(leafs[1], t_depth, 0.02) + R.compute_2dcolor_loss(leafs[0], t_color, None) + R.semantic_2d_ce_loss(leafs[2], label, cw)
    tot, terms = L.losses_2d(raycast_color=leafs2[0], raycast_depth=leafs2[1], raycast_semantic=leafs2[2], images_depth=t_depth,
                             images_color=t_color, target2d_label=label, weight_semantic_class=cw, voxelsize=0.02)
    if torch.isfinite(tot_ref):
        tot_ref.backward(); tot.backward()
        if abs(float(tot.detach()) - float(tot_ref.detach())) > 1e-5 * max(1.0, abs(float(tot_ref.detach()))):
            fail(c, "losses_2d value %g vs %g" % (float(tot), float(tot_ref))); continue
        for a, b, name in zip(leafs2, leafs, ("color", "depth", "semantic")):
            ga, gb = a.grad, b.grad
            fin = torch.isfinite(gb)
            if not torch.allclose(ga[fin], gb[fin], rtol=1e-4, atol=1e-7):
                fail(c, "losses_2d grad %s max err %g" % (name, float((ga[fin] - gb[fin]).abs().max()))); break
    # ---- per-voxel normals (loss.py:285-306)
    if locs.shape[0] > 0 and min(dz, dy, dx) >= 3 and int(locs[-1, 3]) == B - 1:
        vals_sdf = (torch.randn(locs.shape[0], 1, generator=g) * 1.5).to(dev)
        tr = None
        if rng.random() < 0.7:
            q, _ = np.linalg.qr(rng.standard_normal((B, 3, 3)))
            m = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1)); m[:, :3, :3] = q; m[:, :3, 3] = rng.standard_normal((B, 3))
            tr = torch.from_numpy(m.astype(np.float32)).to(dev)
        wn = torch.randn(locs.shape[0], 3, generator=g).to(dev)
        a = vals_sdf.clone().requires_grad_(True); b2 = vals_sdf.clone().requires_grad_(True)
        want_n = R.compute_normals_sparse(locs, a, (dz, dy, dx), tr)
        got_n = compute_normals_sparse(locs, b2, (dz, dy, dx), tr, num_chunks=B)
        (want_n * wn).sum().backward(); (got_n * wn).sum().backward()
        scale = float(a.grad.abs().max()) + 1e-12
        if float((got_n - want_n).abs().max()) > 1e-5 or float((b2.grad - a.grad).abs().max()) > 1e-4 * scale:
            # ill-conditioned voxels (|gradient of the sdf| near the normalisation eps) amplify fp32 rounding: judge both
            # against the same expression in float64
            a64 = vals_sdf.double().clone().requires_grad_(True)
            n64 = R.compute_normals_sparse(locs, a64, (dz, dy, dx), None if tr is None else tr.double())
            (n64 * wn.double()).sum().backward()
            e_mine = float((b2.grad.double() - a64.grad).abs().max()); e_lit = float((a.grad.double() - a64.grad).abs().max())
            # (a voxel whose sdf gradient is tiny and axis-aligned has an exactly-zero true component that fp32 turns into
            #  |q| * ulp / |g|: accepted up to 1e-3 of the largest gradient entry, the bar of the atomically accumulated grads)
            if e_mine > max(4.0 * e_lit, 1e-3 * scale, 1e-5):
                err = (b2.grad.double() - a64.grad).abs()[:, 0]
                top = torch.topk(err, min(3, err.numel()))[1]
                for j in top.tolist():
                    print("   voxel %d loc %s dims %s mine %.9g literal %.9g f64 %.9g  tr %s" % (
                        j, locs[j].tolist(), (dz, dy, dx), float(b2.grad[j]), float(a.grad[j]), float(a64.grad[j]), tr is not None), flush=True)
                fail(c, "normals: value err %g grad err %g (scale %g); vs float64: mine %g literal %g" % (
                    float((got_n - want_n).abs().max()), float((b2.grad - a.grad).abs().max()), scale, e_mine, e_lit))
print("%d cases, %d mismatching, %.1f s" % (cases, bad, time.time() - t0))
sys.exit(1 if bad else 0)

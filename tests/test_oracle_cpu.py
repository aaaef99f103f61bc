"""Known-answer tests of the CPU oracle (oracle/raycast_oracle.c), derivable from the reference code alone
(SURVEY.md section 4): these run without a GPU and pin the restatement's semantics."""
import numpy as np
import pytest

from oracle import oracle as O
from spsg_b200 import synthetic as S


def plane_chunk(dims_zyx=(64, 32, 32), z0=20.3):
    dz, dy, dx = dims_zyx
    z = np.arange(dz, dtype=np.float32)[:, None, None] + np.zeros((dz, dy, dx), np.float32)
    sdf = np.clip(z - z0, -3, 3).astype(np.float32)
    mask = np.abs(sdf) < 3
    locs = np.argwhere(mask).astype(np.int64)
    locs = np.concatenate([locs, np.zeros((locs.shape[0], 1), np.int64)], 1)
    n = locs.shape[0]
    rng = np.random.default_rng(0)
    return dict(locs=locs, sdf=sdf[mask].reshape(n, 1), color=rng.random((n, 3), dtype=np.float32),
                normal=np.tile(np.array([[0, 0, -1]], np.float32), (n, 1)),
                semantic=rng.standard_normal((n, 14)).astype(np.float32)), dims_zyx


def render(chunk, dims, view, intr, w, h, views=1, num_chunks=1, max_pix=64, threads=1, inc=S.RAY_INCREMENT):
    p = O.make_params(dims, w, h, 5.0, 300.0, S.THRESH_SAMPLE_DIST, inc, num_chunks, views, max_pix,
                      chunk["locs"].shape[0])
    sm = O.build_index(chunk["locs"], num_chunks, dims)
    out = O.raycast_forward(p, sm, chunk["sdf"], chunk["color"], chunk["normal"], chunk["semantic"], view, intr,
                            threads=threads)
    return p, sm, out


def test_fronto_parallel_plane_depth():
    chunk, dims = plane_chunk()
    w, h = 48, 40
    view = S.look_at((16.0, 16.0, -30.0), (16.0, 16.0, 0.0), up=(0.0, 1.0, 0.0))[None]
    intr = np.array([[120.0, 120.0, 23.5, 19.5]], np.float32)
    p, sm, out = render(chunk, dims, view, intr, w, h)
    hit = np.isfinite(out["depth"][0])
    assert hit[h // 2, w // 2]
    assert hit.mean() > 0.5
    np.testing.assert_allclose(out["depth"][0][hit], 50.3, atol=2e-3)   # z0 - cam_z
    # hit voxel is the nearest voxel of the crossing: z index 20
    assert set(np.unique(chunk["locs"][out["hit_index"][0][hit], 0])) == {20}
    assert out["mapping3dto2d_num"].sum() == hit.sum()
    # colour is the nearest voxel's payload, not interpolated (kernel.cu:129)
    idx = out["hit_index"][0][hit]
    np.testing.assert_array_equal(out["color"][0][hit], chunk["color"][idx])
    np.testing.assert_array_equal(out["semantic"][0][hit], chunk["semantic"][idx])


def test_empty_chunk_renders_minus_inf():
    dims = (16, 16, 16)
    chunk = dict(locs=np.zeros((0, 4), np.int64), sdf=np.zeros((0, 1), np.float32), color=np.zeros((0, 3), np.float32),
                 normal=np.zeros((0, 3), np.float32), semantic=np.zeros((0, 14), np.float32))
    view = S.look_at((8.0, 8.0, -20.0), (8.0, 8.0, 0.0), up=(0.0, 1.0, 0.0))[None]
    intr = np.array([[50.0, 50.0, 15.5, 11.5]], np.float32)
    _, _, out = render(chunk, dims, view, intr, 32, 24)
    for k in ("color", "depth", "normal", "semantic"):
        assert np.all(np.isneginf(out[k]))
    assert np.all(out["hit_index"] == -1)


def test_backward_of_ones_is_one_per_hit_voxel():
    chunk, dims = plane_chunk()
    w, h = 48, 40
    view = S.look_at((16.0, 16.0, -30.0), (16.0, 16.0, 0.0), up=(0.0, 1.0, 0.0))[None]
    intr = np.array([[120.0, 120.0, 23.5, 19.5]], np.float32)
    p, sm, out = render(chunk, dims, view, intr, w, h)
    g = [np.ones_like(out[k]) for k in ("color", "depth", "normal", "semantic")]
    d = O.raycast_backward(p, *g, sm, out["mapping3dto2d"], out["mapping3dto2d_num"])
    hv = out["mapping3dto2d_num"] > 0
    assert hv.any() and (~hv).any()
    for x in d:
        np.testing.assert_allclose(x[hv], 1.0, atol=1e-5)
        assert np.all(x[~hv] == 0.0)


def test_pixel_cap_keeps_first_pixels_and_means_over_cap():
    chunk, dims = plane_chunk()
    w, h = 48, 40
    view = S.look_at((16.0, 16.0, -30.0), (16.0, 16.0, 0.0), up=(0.0, 1.0, 0.0))[None]
    intr = np.array([[120.0, 120.0, 23.5, 19.5]], np.float32)   # ~6 pixels per voxel side -> > 4 pixels/voxel
    p, sm, out = render(chunk, dims, view, intr, w, h, max_pix=4)
    num = out["mapping3dto2d_num"]
    assert num.max() > 4
    g = [np.ones_like(out[k]) for k in ("color", "depth", "normal", "semantic")]
    d = O.raycast_backward(p, *g, sm, out["mapping3dto2d"], num)
    np.testing.assert_allclose(d[1][num > 0], 1.0, atol=1e-5)   # mean over min(num, cap) kept pixels of ones


def test_zero_normal_keeps_minus_inf():
    chunk, dims = plane_chunk()
    chunk["normal"][:] = 0.0
    view = S.look_at((16.0, 16.0, -30.0), (16.0, 16.0, 0.0), up=(0.0, 1.0, 0.0))[None]
    intr = np.array([[120.0, 120.0, 23.5, 19.5]], np.float32)
    _, _, out = render(chunk, dims, view, intr, 48, 40)
    hit = np.isfinite(out["depth"])
    assert hit.any()
    assert np.all(np.isneginf(out["normal"]))
    assert np.all(np.isfinite(out["color"][hit]))


def test_multi_view_equals_separate_calls_and_threads_are_deterministic():
    batch = S.make_batch([1, 2])
    w, h = 64, 48
    view, intr = S.make_views(2, 2, seed=1)
    intr = (intr * np.array([w / S.WIDTH, h / S.HEIGHT, w / S.WIDTH, h / S.HEIGHT], np.float32)).astype(np.float32)
    p2, sm, out2 = render(batch, S.DIMS_ZYX, view, intr, w, h, views=2, num_chunks=2, threads=4)
    _, _, out2b = render(batch, S.DIMS_ZYX, view, intr, w, h, views=2, num_chunks=2, threads=1)
    np.testing.assert_array_equal(out2["depth"], out2b["depth"])
    np.testing.assert_array_equal(out2["hit_index"], out2b["hit_index"])
    n = batch["locs"].shape[0]
    grads = [np.random.default_rng(3).standard_normal(out2b[k].shape).astype(np.float32)
             for k in ("color", "depth", "normal", "semantic")]
    d2 = O.raycast_backward(p2, *grads, sm, out2b["mapping3dto2d"], out2b["mapping3dto2d_num"])
    acc = [np.zeros_like(x) for x in d2]
    for f in range(2):
        sel = np.array([0, 1]) * 2 + f
        p1, _, out1 = render(batch, S.DIMS_ZYX, view[sel], intr[sel], w, h, views=1, num_chunks=2)
        np.testing.assert_array_equal(out1["depth"], out2b["depth"][sel])
        np.testing.assert_array_equal(out1["mapping3dto2d_num"][:n], out2b["mapping3dto2d_num"][f * n:(f + 1) * n])
        d1 = O.raycast_backward(p1, *[g[sel] for g in grads], sm, out1["mapping3dto2d"], out1["mapping3dto2d_num"])
        for a, x in zip(acc, d1):
            a += x
    for a, x in zip(acc, d2):
        np.testing.assert_allclose(a, x, rtol=1e-5, atol=1e-6)


def test_occ_render_matches_hit_mask_of_thin_shell():
    sdf, _ = S.sdf_volume(0)
    occ = (np.abs(sdf) < 1.0).astype(np.uint8)[None, None]
    w, h = 64, 48
    view, intr = S.make_views(1, 1)
    intr = (intr * np.array([w / S.WIDTH, h / S.HEIGHT, w / S.WIDTH, h / S.HEIGHT], np.float32)).astype(np.float32)
    p = O.make_params(S.DIMS_ZYX, w, h, S.DEPTH_MIN, S.DEPTH_MAX, 0, S.RAY_INCREMENT, 1)
    got = O.raycast_occ(p, occ, view, intr)
    assert got.shape == (1, 1, h, w) and set(np.unique(got)) <= {0, 1}
    batch = S.make_batch([0])
    _, _, out = render(batch, S.DIMS_ZYX, view, intr, w, h)
    hit = np.isfinite(out["depth"][0])
    # every ray that finds the SDF zero crossing passes through the |sdf| < 1 shell with 0.9-voxel steps
    assert (got[0, 0][hit] == 1).mean() > 0.98


def test_golden_vectors_from_reference_extension():
    """tests/golden/*.npz hold outputs of the *reference* CUDA extension (made by tests/golden/make_golden.py on a
    B200).  The CPU restatement must reproduce them: same hit mask up to eps-ambiguous rays, values to 1e-4."""
    import glob
    import os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sphere_*.npz")))
    if not files:
        pytest.skip("no golden vectors committed yet")
    for f in files:
        g = np.load(f)
        dims = tuple(int(x) for x in g["dims_zyx"])
        h, w = g["depth"].shape[1:]
        p = O.make_params(dims, w, h, float(g["depth_min"]), float(g["depth_max"]), float(g["thresh"]),
                          float(g["inc"]), int(g["num_chunks"]), 1, 64, g["locs"].shape[0])
        sm = O.build_index(g["locs"], int(g["num_chunks"]), dims)
        out = O.raycast_forward(p, sm, g["sdf"], g["color"], g["normal"], g["semantic"], g["view"], g["intr"],
                                threads=O.max_threads())
        hit_ref, hit = np.isfinite(g["depth"]), np.isfinite(out["depth"])
        assert (hit_ref != hit).sum() <= max(2, hit.size // 5000), f
        both = hit_ref & hit
        np.testing.assert_allclose(out["depth"][both], g["depth"][both], rtol=1e-4)
        same = (out["color"][both] == g["color_img"][both]).all(-1)
        assert same.mean() > 0.999
        d = O.raycast_backward(p, g["g_color"], g["g_depth"], g["g_normal"], g["g_semantic"], sm,
                               out["mapping3dto2d"], out["mapping3dto2d_num"])
        if (hit_ref != hit).sum() == 0:
            np.testing.assert_allclose(d[1], g["d_depth"], rtol=1e-3, atol=1e-5)
            np.testing.assert_allclose(d[3], g["d_semantic"], rtol=1e-3, atol=1e-5)


def test_loss_restatements_match_reference_python_golden():
    """oracle/losses_ref.py against outputs of the reference's own loss.py (tests/golden/make_golden_losses.py)."""
    import os
    import torch
    from oracle import losses_ref as R
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses_ref.npz"))
    locs, sdf = torch.from_numpy(g["n_locs"]), torch.from_numpy(g["n_sdf"])
    dims, w = tuple(int(v) for v in g["n_dims"]), torch.from_numpy(g["n_w"])
    for name, tr in (("normals_t", torch.from_numpy(g["n_transform"])), ("normals_id", None)):
        v = sdf.clone().requires_grad_(True)
        n = R.compute_normals_sparse(locs, v, dims, tr)
        (n * w).sum().backward()
        assert np.abs(n.detach().numpy() - g[name]).max() < 1e-6
        assert np.abs(v.grad.numpy() - g[name + "_dsdf"]).max() < 1e-4 * np.abs(g[name + "_dsdf"]).max()
    pred, tgt = torch.from_numpy(g["c_pred"]), torch.from_numpy(g["c_tgt"])
    for name, ww in (("color_w", torch.from_numpy(g["c_weight"])), ("color_now", None)):
        p = pred.clone().requires_grad_(True)
        l = R.compute_2dcolor_loss(p, tgt, ww)
        l.backward()
        assert abs(float(l) - float(g[name])) < 1e-6
        hole = torch.isinf(pred)
        gr = p.grad.clone()
        gr[hole] = 0
        assert np.abs(gr.numpy() - g[name + "_grad"]).max() < 1e-7

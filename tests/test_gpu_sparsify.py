"""Producer glue (sparsify.py, csrc/spsg_sparsify.cu) against the reference's own PyTorch expressions
(torch/train.py:494-509): identical voxel rows in identical order, identical values, identical gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _literal_locs(sdf, truncation, empty=None):
    mask = torch.abs(sdf.detach()[:, 0]) < truncation           # train.py:495-497
    if empty is not None:
        mask = mask & ~empty[:, 0]
    locs = torch.nonzero(mask)
    return torch.cat([locs[:, 1:], locs[:, :1]], 1)             # train.py:498


def _literal_gather(head, locs):
    return head[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]   # train.py:499


def _volume(shape, device, seed, channels=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape[0], channels, *shape[1:], generator=g) * 3.0).to(device)


@pytest.mark.parametrize("shape", [(8, 128, 64, 64), (2, 5, 7, 9), (1, 33, 31, 65), (3, 16, 16, 16)])
@pytest.mark.parametrize("use_empty", [False, True])
def test_locs_match_nonzero_order(cuda_device, shape, use_empty):
    from spsg_b200 import sparsify
    sdf = _volume(shape, cuda_device, 1)
    sdf.view(-1)[::97] = float("nan")                           # NaN never passes |x| < t
    sdf.view(-1)[5::131] = 3.0                                  # exactly the truncation: excluded
    empty = (_volume(shape, cuda_device, 2) > 1.0) if use_empty else None
    got = sparsify.sparse_locs(sdf, 3.0, empty)
    want = _literal_locs(sdf, 3.0, empty)
    assert got.dtype == torch.int64 and got.shape == want.shape
    assert torch.equal(got, want)
    assert got.shape[0] > 0


def test_locs_edge_cases(cuda_device):
    from spsg_b200 import sparsify
    shape = (2, 8, 8, 8)
    sdf = _volume(shape, cuda_device, 3)
    assert sparsify.sparse_locs(sdf, 0.0).shape == (0, 4)                         # nothing passes
    every = sparsify.sparse_locs(sdf, 1e9)
    assert torch.equal(every, _literal_locs(sdf, 1e9)) and every.shape[0] == sdf.numel()
    off = torch.zeros(sdf.numel() + 3, device=cuda_device)                         # a view that is not 16-byte aligned
    off[3:] = sdf.reshape(-1)
    view = off[3:].view(sdf.shape)
    assert torch.equal(sparsify.sparse_locs(view, 2.0), _literal_locs(sdf, 2.0))
    assert torch.equal(sparsify.sparse_locs(sdf[:, 0], 2.0), _literal_locs(sdf, 2.0))   # (B,Dz,Dy,Dx) input
    with pytest.raises(RuntimeError):
        sparsify.sparse_locs(sdf.cpu(), 2.0)


@pytest.mark.parametrize("shape", [(2, 24, 16, 20), (1, 5, 7, 9)])
def test_gather_values_and_gradients(cuda_device, shape):
    from spsg_b200 import sparsify
    sdf = _volume(shape, cuda_device, 4)
    heads = [_volume(shape, cuda_device, 10 + c, channels=c) for c in (3, 14, 20, 1, 2)]   # > 4 heads, > 16 channels
    locs = _literal_locs(sdf, 2.5)
    mine_in = [t.clone().requires_grad_(True) for t in [sdf] + heads]
    ref_in = [t.clone().requires_grad_(True) for t in [sdf] + heads]
    got = sparsify.gather_dense(locs, *mine_in)
    want = [_literal_gather(t, locs) for t in ref_in]
    g = torch.Generator(device="cpu").manual_seed(9)
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)
    ws = [torch.randn(b.shape, generator=g).to(cuda_device) for b in want]
    sum((a * w).sum() for a, w in zip(got, ws)).backward()
    sum((b * w).sum() for b, w in zip(want, ws)).backward()
    for a, b in zip(mine_in, ref_in):
        assert torch.equal(a.grad, b.grad)


def test_sparsify_predictions_feeds_the_raycaster(cuda_device):
    """train.py:494-509 in one call, on the synthetic chunk laid out as dense heads."""
    from spsg_b200 import sparsify, synthetic as S
    batch = S.make_batch([0, 1])
    B = 2
    dz, dy, dx = S.DIMS_ZYX
    locs_np, sdf_np = batch["locs"], batch["sdf"]
    dense_sdf = torch.full((B, 1, dz, dy, dx), 10.0)
    l = torch.from_numpy(locs_np)
    dense_sdf[l[:, 3], 0, l[:, 0], l[:, 1], l[:, 2]] = torch.from_numpy(sdf_np[:, 0])
    dense_sem = torch.randn(B, 14, dz, dy, dx)
    dense_sdf, dense_sem = dense_sdf.to(cuda_device), dense_sem.to(cuda_device)
    locs, vals_sdf, vals_sem = sparsify.sparsify_predictions(dense_sdf, S.TRUNCATION, None, dense_sem)
    assert torch.equal(locs.cpu(), l)                                  # the generator's own (nonzero-ordered) voxel list
    assert torch.equal(vals_sdf.cpu(), torch.from_numpy(sdf_np))
    assert torch.equal(vals_sem, _literal_gather(dense_sem, locs))


def _dense_heads(device, seeds):
    """the synthetic chunks laid out as the generator's dense heads (SDF, colour, 14 logits)"""
    from spsg_b200 import synthetic as S
    batch = S.make_batch(seeds)
    B = len(seeds)
    dz, dy, dx = S.DIMS_ZYX
    l = torch.from_numpy(batch["locs"])
    idx = (l[:, 3], l[:, 0], l[:, 1], l[:, 2])
    sdf = torch.full((B, dz, dy, dx), 2.0 * S.TRUNCATION)
    sdf[idx] = torch.from_numpy(batch["sdf"][:, 0])
    col = torch.zeros(B, dz, dy, dx, 3)
    col[idx] = torch.from_numpy(batch["color"])
    sem = torch.zeros(B, dz, dy, dx, 14)
    sem[idx] = torch.from_numpy(batch["semantic"])
    return (sdf.unsqueeze(1).contiguous().to(device), col.permute(0, 4, 1, 2, 3).contiguous().to(device),
            sem.permute(0, 4, 1, 2, 3).contiguous().to(device))


def _raycaster(B, F, device, n_max):
    from spsg_b200 import synthetic as S
    from spsg_b200.raycast_rgbd import RaycastRGBD
    return RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                       max_num_frames=F, max_num_locs_per_sample=n_max, device=device)


@pytest.mark.parametrize("fused", [False, True])
def test_direct_feed_equals_the_locs_round_trip(cuda_device, fused):
    """sparsify_predictions(..., raycaster=m) writes m's voxel index and dense SDF brick in the pass that writes locs; the
    forward over those rows then skips its fill + index passes (SPSG_FLAG_INDEX_PREBUILT).  Renderings, index, counters and
    voxel gradients must equal, bit for bit, what the plain route (locs -> fill -> index) gives on a second module."""
    from spsg_b200 import _native as N, losses, normals, sparsify, synthetic as S
    from tests.common import views
    B, F = 2, 2
    sdf, col, sem = _dense_heads(cuda_device, [0, 1])
    _, _, view, intr = views(B, F, cuda_device, seed=3)
    g = torch.Generator(device="cpu").manual_seed(1)
    t_depth = (torch.rand(B * F, S.HEIGHT, S.WIDTH, generator=g) * 3.0).to(cuda_device)
    t_color = torch.rand(B * F, S.HEIGHT, S.WIDTH, 3, generator=g).to(cuda_device)
    t_label = torch.randint(0, 15, (B * F, S.HEIGHT, S.WIDTH), generator=g).to(torch.uint8).to(cuda_device)
    out = []
    for feed in (True, False):
        m = _raycaster(B, F, cuda_device, 200000)
        heads = [t.detach().clone().requires_grad_(True) for t in (sdf, col, sem)]
        locs, v_sdf, v_col, v_sem = sparsify.sparsify_predictions(heads[0], S.TRUNCATION, None, heads[1], heads[2],
                                                                  raycaster=m if feed else None)
        assert (m.workspace._prebuilt is not None) == feed
        if feed:
            index_before = m.sparse_mapping.clone()
        nrm = normals.compute_normals_sparse(locs, v_sdf.detach(), S.DIMS_ZYX, transform=torch.inverse(view[::F]), num_chunks=B)
        if fused:
            total, _, _ = losses.render_with_2d_losses(m, locs, v_sdf, v_col, nrm, v_sem, view, intr, images_depth=t_depth,
                                                       images_color=t_color, target2d_label=t_label, voxelsize=S.VOXELSIZE)
        else:
            c, d, nn, s = m(locs, v_sdf, v_col, nrm, v_sem, view, intr)
            hit = d != -float("inf")
            total = d[hit].sum() * 0.01 + c[hit].sum() + s[hit].sum() * 0.1
        if feed:
            assert m.workspace._prebuilt is not None          # taken by the forward and still valid after it
            assert torch.equal(index_before, m.sparse_mapping)  # the forward did not rebuild the index
        total.backward()
        out.append((locs, m.sparse_mapping.clone(), m.image_color.clone(), m.image_depth.clone(), m.image_normal.clone(),
                    m.image_semantic.clone(), m.mapping3dto2d_num[:locs.shape[0] * F].clone(), total.detach().clone(),
                    heads[0].grad.clone(), heads[1].grad.clone(), heads[2].grad.clone()))
    names = ("locs", "sparse_mapping", "color", "depth", "normal", "semantic", "mapping3dto2d_num", "loss", "d_sdf", "d_color",
             "d_semantic")
    for a, b, name in zip(out[0], out[1], names):
        if name.startswith("d_") or name == "loss":   # gradients: per-voxel means of atomically registered pixels
            torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-6, msg=name)
        else:
            assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a,
                               b.view(torch.int32) if b.dtype == torch.float32 else b), name
    assert int((out[0][3] != -float("inf")).sum()) > 1000


def test_direct_feed_mark_is_dropped_by_other_renders(cuda_device):
    """The prebuilt index + brick belong to one locs tensor.  A render of other rows on the same module overwrites them:
    the mark must be gone, and a later render of the first rows must rebuild (and still be right)."""
    from spsg_b200 import sparsify, synthetic as S
    from tests.common import scene_tensors, views
    sdf, col, sem = _dense_heads(cuda_device, [0])
    _, _, view, intr = views(1, 1, cuda_device, seed=5)
    m = _raycaster(1, 1, cuda_device, 200000)
    locs, v_sdf, v_col, v_sem = sparsify.sparsify_predictions(sdf, S.TRUNCATION, None, col, sem, raycaster=m)
    nrm = torch.zeros(locs.shape[0], 3, device=cuda_device)
    nrm[:, 2] = 1.0
    with torch.no_grad():
        first = [t.clone() for t in m(locs, v_sdf, v_col, nrm, v_sem, view, intr)]
        assert m.workspace._prebuilt is not None
        _, other = scene_tensors([7], cuda_device)
        m(other["locs"], other["sdf"], other["color"], other["normal"], other["semantic"], view, intr)
        assert m.workspace._prebuilt is None
        again = m(locs, v_sdf, v_col, nrm, v_sem, view, intr)
        for a, b in zip(first, again):
            assert torch.equal(a.view(torch.int32), b.view(torch.int32))
        # a modified locs tensor no longer matches its mark either
        locs2, v2, c2, s2 = sparsify.sparsify_predictions(sdf, S.TRUNCATION, None, col, sem, raycaster=m)
        locs2[0, 0] += 0
        assert m.workspace.take_prebuilt(locs2, 0) == 0


def test_counted_locs_split(cuda_device):
    """count_locs now, rows later (after other work): same rows; a count of other tensors is refused."""
    from spsg_b200 import sparsify
    sdf = _volume((2, 24, 16, 20), cuda_device, 11)
    empty = _volume((2, 24, 16, 20), cuda_device, 12) > 1.0
    counted = sparsify.count_locs(sdf, 2.5, empty)
    other = sparsify.sparse_locs(_volume((2, 24, 16, 20), cuda_device, 13), 1.0)      # unrelated compaction in between
    want = _literal_locs(sdf, 2.5, empty)
    assert counted.n == want.shape[0] and other.shape[0] != counted.n
    assert torch.equal(sparsify.sparse_locs(sdf, 2.5, empty, counted=counted), want)
    with pytest.raises(RuntimeError):
        sparsify.sparse_locs(sdf, 2.0, empty, counted=counted)                        # another truncation
    with pytest.raises(RuntimeError):
        sparsify.sparse_locs(sdf.clone(), 2.5, empty, counted=counted)                # another tensor
    sdf[0, 0, 0, 0, 0] = 0.0
    with pytest.raises(RuntimeError):
        sparsify.sparse_locs(sdf, 2.5, empty, counted=counted)                        # modified since the count

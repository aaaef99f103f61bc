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

"""spsg_b200.normals.compute_normals_sparse (fused gather kernels through the C ABI) against the literal PyTorch
restatement of the reference's loss.compute_normals_sparse (oracle/losses_ref.py) with autograd.
Tolerances: values 1e-5 absolute (unit vectors; the 3x3 products are summed in a different order), SDF gradients 1e-4 of
the largest gradient."""
import numpy as np
import pytest
import torch

from tests.common import scene_tensors, views

pytestmark = pytest.mark.gpu


def _inputs(device, seeds, dims=None):
    from spsg_b200 import synthetic as S
    dims = S.DIMS_ZYX if dims is None else dims
    _, t = scene_tensors(seeds, device, dims_zyx=dims)
    view, _, _, _ = views(len(seeds), 1, device, seed=3)
    transform = torch.inverse(torch.from_numpy(view).to(device))       # train.py:544
    return t["locs"], t["sdf"], transform.contiguous(), dims


@pytest.mark.parametrize("with_transform", [True, False])
def test_normals_match_reference_expression(cuda_device, with_transform):
    from oracle import losses_ref as R
    from spsg_b200.normals import compute_normals_sparse
    locs, sdf, transform, dims = _inputs(cuda_device, [31, 32])
    tr = transform if with_transform else None
    torch.manual_seed(0)
    w = torch.randn(locs.shape[0], 3, device=cuda_device)

    a = sdf.clone().requires_grad_(True)
    want = R.compute_normals_sparse(locs, a, dims, tr)
    (want * w).sum().backward()

    b = sdf.clone().requires_grad_(True)
    got = compute_normals_sparse(locs, b, dims, tr, num_chunks=2)
    (got * w).sum().backward()

    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-5
    # border voxels and isolated voxels: zero normal (the raycaster then leaves the pixel's normal at -inf, kernel.cu:220)
    zero = (want == 0).all(dim=1)
    assert bool((got[zero] == 0).all())
    scale = float(a.grad.abs().max())
    assert scale > 0
    assert float((b.grad - a.grad).abs().max()) < 1e-4 * scale


def test_normals_small_grid_with_border_voxels(cuda_device):
    """Every voxel of a tiny grid present: exercises the border rule on all six faces and the eps branch of normalize
    (constant regions have zero gradient)."""
    from oracle import losses_ref as R
    from spsg_b200.normals import compute_normals_sparse
    dims = (6, 5, 7)
    g = torch.Generator().manual_seed(3)
    vol = torch.randn(2, *dims, generator=g)
    vol[1, 2:4, 1:4, 2:5] = 0.25           # flat patch -> |m| = 0 -> eps branch
    locs = torch.nonzero(torch.ones(2, *dims, dtype=torch.bool))
    locs = torch.cat([locs[:, 1:], locs[:, :1]], 1).contiguous().to(cuda_device)
    sdf = vol[locs[:, 3].cpu(), locs[:, 0].cpu(), locs[:, 1].cpu(), locs[:, 2].cpu()].reshape(-1, 1).to(cuda_device)
    transform = torch.eye(4).repeat(2, 1, 1)
    transform[1, :3, :3] = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    transform = transform.to(cuda_device)
    w = torch.randn(locs.shape[0], 3, generator=g).to(cuda_device)
    a = sdf.clone().requires_grad_(True)
    want = R.compute_normals_sparse(locs, a, dims, transform)
    (want * w).sum().backward()
    b = sdf.clone().requires_grad_(True)
    got = compute_normals_sparse(locs, b, dims, transform)
    (got * w).sum().backward()
    assert float((got - want).abs().max()) < 1e-5
    assert float((b.grad - a.grad).abs().max()) < 1e-4 * float(a.grad.abs().max())


def test_normals_feed_the_raycaster(cuda_device):
    """End of the producer chain (train.py:542 -> :626): normals computed here render bit-identically to normals computed
    by the reference expression wherever the two agree bitwise, and the second gradient path reaches the SDF."""
    from spsg_b200 import synthetic as S
    from spsg_b200.normals import compute_normals_sparse
    from spsg_b200.raycast_rgbd import RaycastRGBD
    _, t = scene_tensors([33], cuda_device)
    view_np, intr_np, view, intr = views(1, 1, cuda_device, seed=1, width=96, height=64)
    sdf = t["sdf"].clone().requires_grad_(True)
    normals = compute_normals_sparse(t["locs"], sdf, S.DIMS_ZYX, torch.inverse(view), num_chunks=1)
    rcst = RaycastRGBD(1, S.DIMS_ZYX, 96, 64, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                       max_num_locs_per_sample=t["locs"].shape[0], device=cuda_device)
    color, depth, normal_img, sem = rcst(t["locs"], sdf, t["color"], normals, t["semantic"], view, intr)
    hit = depth != -float("inf")
    assert hit.any()
    seen = normal_img[hit]
    lens = seen[seen[:, 0] != -float("inf")].norm(dim=1)
    assert float((lens - 1).abs().max()) < 1e-4            # unit normals came through
    (normal_img[hit][:, 2].clamp(min=-2).sum() + depth[hit].sum() * 0.01).backward()
    assert float(sdf.grad.abs().sum()) > 0


def test_normals_argument_errors(cuda_device):
    from spsg_b200.normals import compute_normals_sparse
    with pytest.raises(RuntimeError, match="no CPU path"):
        compute_normals_sparse(torch.zeros(1, 4, dtype=torch.long), torch.zeros(1, 1), (4, 4, 4))
    out = compute_normals_sparse(torch.zeros(0, 4, dtype=torch.long, device=cuda_device), torch.zeros(0, 1, device=cuda_device), (4, 4, 4))
    assert out.shape == (0, 3)


def test_normals_match_reference_python_golden(cuda_device):
    """Against outputs of the reference's own loss.compute_normals_sparse (tests/golden/losses_ref.npz)."""
    import os
    from spsg_b200.normals import compute_normals_sparse
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses_ref.npz"))
    locs = torch.from_numpy(g["n_locs"]).to(cuda_device)
    dims, w = tuple(int(v) for v in g["n_dims"]), torch.from_numpy(g["n_w"]).to(cuda_device)
    for name, tr in (("normals_t", torch.from_numpy(g["n_transform"]).to(cuda_device)), ("normals_id", None)):
        v = torch.from_numpy(g["n_sdf"]).to(cuda_device).requires_grad_(True)
        n = compute_normals_sparse(locs, v, dims, tr, num_chunks=2)
        (n * w).sum().backward()
        assert np.abs(n.detach().cpu().numpy() - g[name]).max() < 1e-5
        want = g[name + "_dsdf"]
        assert np.abs(v.grad.cpu().numpy() - want).max() < 1e-4 * np.abs(want).max()

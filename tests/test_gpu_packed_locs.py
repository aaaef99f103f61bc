"""Host-fed calls: the reference's int64 voxel rows packed on the host into uint32 cell indices (spsg_pack_locs_host) and
rendered with SPSG_FLAG_PACKED_LOCS must give exactly what the int64 rows give."""
import numpy as np
import pytest
import torch

from spsg_b200 import synthetic as S


def test_pack_locs_host_cpu():
    """(no GPU needed) cell = ((b*Dz+z)*Dy+y)*Dx+x; rows outside the grid give the sentinel; any thread count."""
    from spsg_b200 import raycast_rgbd_cuda as rc
    rng = np.random.default_rng(3)
    dims, B = (7, 5, 9), 3
    locs = np.stack([rng.integers(0, dims[0], 1000), rng.integers(0, dims[1], 1000), rng.integers(0, dims[2], 1000),
                     rng.integers(0, B, 1000)], 1).astype(np.int64)
    locs[5] = (-1, 0, 0, 0); locs[6] = (0, 5, 0, 0); locs[7] = (0, 0, 0, 3); locs[8] = (1 << 40, 0, 0, 0)
    want = ((locs[:, 3] * dims[0] + locs[:, 0]) * dims[1] + locs[:, 1]) * dims[2] + locs[:, 2]
    want[5:9] = 0xffffffff
    for threads in (1, 3, 64):
        got = rc.pack_locs_host(torch.from_numpy(locs), B, dims, threads=threads)
        assert got.dtype == torch.int32 and got.shape == (1000,)
        assert np.array_equal(got.numpy().view(np.uint32), want.astype(np.uint32))
    with pytest.raises(RuntimeError):
        rc.pack_locs_host(torch.from_numpy(locs).int(), B, dims)
    assert rc.pack_locs_host(torch.zeros(0, 4, dtype=torch.int64), B, dims).shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_packed_locs_render_like_int64_rows(cuda_device, fused):
    from spsg_b200 import losses, raycast_rgbd_cuda as rc
    from spsg_b200.raycast_rgbd import RaycastRGBD
    from tests.common import scene_tensors, views
    B, F = 2, 2
    batch, t = scene_tensors([3, 4], cuda_device)
    _, _, view, intr = views(B, F, cuda_device, seed=2)
    cells = rc.pack_locs_host(torch.from_numpy(batch["locs"]), B, S.DIMS_ZYX).to(cuda_device)
    g = torch.Generator(device="cpu").manual_seed(5)
    t_depth = (torch.rand(B * F, S.HEIGHT, S.WIDTH, generator=g) * 3.0).to(cuda_device)
    t_label = torch.randint(0, 15, (B * F, S.HEIGHT, S.WIDTH), generator=g).to(torch.uint8).to(cuda_device)
    out = []
    for locs in (t["locs"], cells):
        m = RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                        max_num_frames=F, max_num_locs_per_sample=200000, device=cuda_device)
        sdf = t["sdf"].detach().clone().requires_grad_(True)
        sem = t["semantic"].detach().clone().requires_grad_(True)
        if fused:
            total, _, _ = losses.render_with_2d_losses(m, locs, sdf, t["color"], t["normal"], sem, view, intr,
                                                       images_depth=t_depth, target2d_label=t_label, voxelsize=S.VOXELSIZE)
        else:
            c, d, n, s = m(locs, sdf, t["color"], t["normal"], sem, view, intr)
            hit = d != -float("inf")
            total = d[hit].sum() * 0.01 + s[hit].sum() * 0.1
        total.backward()
        out.append((m.sparse_mapping.clone(), m.image_depth.clone(), m.image_semantic.clone(), m.image_color.clone(),
                    total.detach().clone(), sdf.grad.clone(), sem.grad.clone()))
    for a, b, name in zip(out[0], out[1], ("sparse_mapping", "depth", "semantic", "color", "loss", "d_sdf", "d_semantic")):
        if name in ("loss", "d_sdf", "d_semantic"):
            torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-6, msg=name)
        else:
            assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a,
                               b.view(torch.int32) if b.dtype == torch.float32 else b), name
    assert int((out[0][1] != -float("inf")).sum()) > 1000

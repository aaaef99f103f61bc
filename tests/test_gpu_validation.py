"""Argument validation of the raw-pointer entry points: wrong dtypes / short tensors / a backward
without its forward raise instead of reading out of bounds; voxel rows outside the grid are skipped."""
import pytest
import torch

from tests.common import scene_tensors, views

pytestmark = pytest.mark.gpu


def _rc(dev, n, B=1, F=1, w=64, h=48):
    from spsg_b200 import synthetic as S
    from spsg_b200.raycast_rgbd import RaycastRGBD
    return RaycastRGBD(B, S.DIMS_ZYX, w, h, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, max_num_frames=F,
                       max_num_locs_per_sample=n, device=dev)


def test_fused_entry_rejects_bad_inputs(cuda_device):
    from spsg_b200.losses import render_with_2d_losses
    _, t = scene_tensors([1], cuda_device)
    n = t["locs"].shape[0]
    _, _, view, intr = views(1, 1, cuda_device, width=64, height=48)
    rc = _rc(cuda_device, n)
    depth = torch.rand(1, 48, 64, device=cuda_device)
    ok = (t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view, intr)
    render_with_2d_losses(rc, *ok, images_depth=depth)
    bad = [
        (t["locs"].int(), *ok[1:]),                                    # int32 locs
        (ok[0], t["sdf"].half(), *ok[2:]),                             # fp16 payload
        (*ok[:4], t["semantic"][: n // 2].contiguous(), *ok[5:]),      # short semantic rows
        (*ok[:5], view[:, :2].contiguous(), intr),                     # short camera
        (*ok[:2], t["color"].t(), *ok[3:]),                            # non-contiguous
        (*ok[:5], view.cpu(), intr),                                   # host tensor
    ]
    for args in bad:
        with pytest.raises(RuntimeError):
            render_with_2d_losses(rc, *args, images_depth=depth)
    with pytest.raises(RuntimeError):
        render_with_2d_losses(rc, *ok, images_depth=depth.double())


def test_backward_without_matching_forward_raises(cuda_device):
    from spsg_b200 import raycast_rgbd_cuda as native
    from spsg_b200 import synthetic as S
    _, t = scene_tensors([2], cuda_device)
    n = t["locs"].shape[0]
    rc = _rc(cuda_device, n)
    g = [torch.zeros_like(x) for x in (rc.image_color, rc.image_depth, rc.image_normal, rc.image_semantic)]
    dims = [1, 64, 64, 128, n]
    # a module that never ran a forward: neither its own workspace nor the per-buffer cache holds a work list
    with pytest.raises(RuntimeError, match="without a matching forward"):
        native.backward(*g, rc.sparse_mapping, rc.mapping3dto2d, rc.mapping3dto2d_num, dims, rc.d_color, rc.d_depth, rc.d_normal,
                        rc.d_semantic, workspace_owner=rc.workspace)
    with pytest.raises(RuntimeError, match="without a matching forward"):
        native.backward(*g, rc.sparse_mapping, rc.mapping3dto2d, rc.mapping3dto2d_num, dims, rc.d_color, rc.d_depth, rc.d_normal,
                        rc.d_semantic)
    # after a forward with other parameters (fewer voxels) the stamp does not match either
    _, _, view, intr = views(1, 1, cuda_device, width=64, height=48)
    half = n // 2
    rc(t["locs"][:half].contiguous(), t["sdf"][:half].contiguous(), t["color"][:half].contiguous(),
       t["normal"][:half].contiguous(), t["semantic"][:half].contiguous(), view, intr)
    with pytest.raises(RuntimeError, match="without a matching forward"):
        native.backward(*g, rc.sparse_mapping, rc.mapping3dto2d, rc.mapping3dto2d_num, dims, rc.d_color, rc.d_depth, rc.d_normal,
                        rc.d_semantic, workspace_owner=rc.workspace)


def test_voxel_rows_outside_the_grid_are_skipped(cuda_device):
    _, t = scene_tensors([3], cuda_device)
    n = t["locs"].shape[0]
    _, _, view, intr = views(1, 1, cuda_device, width=64, height=48)
    rc = _rc(cuda_device, n + 4)
    want = [o.clone() for o in rc(t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view, intr)]
    junk = torch.tensor([[-1, 0, 0, 0], [0, 64, 0, 0], [0, 0, 0, 1], [128, 3, 3, 0]], device=cuda_device)
    locs = torch.cat([t["locs"], junk]).contiguous()
    pad = lambda x: torch.cat([x, torch.ones(4, x.shape[1], device=cuda_device)]).contiguous()
    got = rc(locs, pad(t["sdf"]), pad(t["color"]), pad(t["normal"]), pad(t["semantic"]), view, intr)
    for a, b in zip(got, want):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_depth2normals_rejects_frames_larger_than_its_buffers(cuda_device):
    from spsg_b200.depth_utils import Depth2Normals
    d2n = Depth2Normals(2, 64, 48, 5.0, 300.0, 4)
    intr = torch.tensor([[60.0, 60.0, 31.5, 23.5]] * 3, device=cuda_device)
    with pytest.raises(RuntimeError):
        d2n(torch.ones(3, 1, 48, 64, device=cuda_device), intr)        # batch too large
    with pytest.raises(RuntimeError):
        d2n(torch.ones(2, 1, 64, 64, device=cuda_device), intr[:2])    # frame too large
    assert d2n(torch.ones(2, 1, 48, 64, device=cuda_device), intr[:2].contiguous()) is not None

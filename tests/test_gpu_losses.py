"""Fused raycast + 2D losses (C ABI spsg_raycast_forward_loss / spsg_raycast_backward_loss) against the literal
PyTorch expressions of the reference (oracle/losses_ref.py) applied to the un-fused rendering with autograd.
Tolerances: loss values 1e-5 relative; voxel gradients 1e-3 relative (sums of a few dozen fp32 terms)."""
import numpy as np
import pytest
import torch

from tests.common import scene_tensors, views

pytestmark = pytest.mark.gpu


def _setup(device, B, F, w, h, weight_color=False):
    from oracle import losses_ref as R
    from spsg_b200 import synthetic as S
    from spsg_b200.raycast_rgbd import RaycastRGBD
    seeds = list(range(20, 20 + B))
    _, pred = scene_tensors(seeds, device, payload="prediction")
    _, tgt = scene_tensors([s + 1000 for s in seeds], device, payload="target")
    n = max(pred["locs"].shape[0], tgt["locs"].shape[0])
    _, _, view, intr = views(B, F, device, seed=2, width=w, height=h)
    rc = RaycastRGBD(B, S.DIMS_ZYX, w, h, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                     max_num_frames=F, max_num_locs_per_sample=n // B + 1000, device=device)
    with torch.no_grad():
        t_color, t_depth, _, t_sem = rc(tgt["locs"], tgt["sdf"], tgt["color"], tgt["normal"], tgt["semantic"], view, intr)
        label = R.labels_from_render(t_sem.clone())                      # (I,H,W,1) uint8
        images_depth = torch.where(t_depth != -float("inf"), t_depth * S.VOXELSIZE, torch.zeros_like(t_depth))
        gen = torch.Generator(device="cpu").manual_seed(5)
        holes = (torch.rand(images_depth.shape, generator=gen) < 0.05).to(device)
        images_depth = images_depth.masked_fill(holes, 0.0).unsqueeze(1).contiguous()   # (I,1,H,W) like train.py:535
        images_color = torch.where(t_color != -float("inf"), t_color, torch.full_like(t_color, 0.5)).contiguous()
    wc = None
    if weight_color:
        wc = torch.ones(B * F, 1, h, w, device=device)
        wc[:, :, : h // 2] = 3.0
    cw = torch.tensor(S.CLASS_WEIGHTS, dtype=torch.float32, device=device)
    return rc, pred, view, intr, images_depth, images_color, wc, label, cw


@pytest.mark.parametrize("B,F,weight_color", [(1, 1, False), (2, 2, True)])
def test_fused_losses_match_literal_pytorch(cuda_device, B, F, weight_color):
    from oracle import losses_ref as R
    from spsg_b200 import synthetic as S
    from spsg_b200.losses import labels_from_render, render_with_2d_losses
    w, h = 160, 128
    rc, pred, view, intr, images_depth, images_color, wc, label, cw = _setup(cuda_device, B, F, w, h, weight_color)
    wd, wcl, ws = 1.0, 0.7, 0.3

    def leafs():
        return [pred[k].clone().requires_grad_(True) for k in ("sdf", "color", "normal", "semantic")]

    # un-fused: module rendering + the reference's literal loss expressions + autograd
    sdf, col, nrm, sem = leafs()
    r_color, r_depth, r_normal, r_sem = rc(pred["locs"], sdf, col, nrm, sem, view, intr)
    l_depth = R.depth_l1_loss(r_depth, images_depth, S.VOXELSIZE)
    l_color = R.compute_2dcolor_loss(r_color, images_color, wc)
    l_sem = R.semantic_2d_ce_loss(r_sem, label, cw)
    total = wd * l_depth + wcl * l_color + ws * l_sem
    total.backward()
    want = [x.grad.clone() for x in (sdf, col, nrm, sem)]
    want_losses = torch.stack([l_depth, l_color, l_sem, total]).detach()
    want_images = [x.detach().clone() for x in (r_color, r_depth, r_normal, r_sem)]
    assert torch.equal(labels_from_render(r_sem.detach()), R.labels_from_render(r_sem.detach())[..., 0])

    # fused
    sdf2, col2, nrm2, sem2 = leafs()
    total2, terms2, images2 = render_with_2d_losses(
        rc, pred["locs"], sdf2, col2, nrm2, sem2, view, intr, images_depth=images_depth, images_color=images_color,
        weight_color=wc, target2d_label=label, weight_semantic_class=cw, voxelsize=S.VOXELSIZE,
        weight_depth_loss=wd, weight_color_loss=wcl, weight_semantic_loss=ws)
    (total2 * 2.0).backward()   # non-unit upstream gradient exercises grad_scale
    got_losses = torch.cat([terms2, total2[None]]).detach()
    torch.testing.assert_close(got_losses, want_losses, rtol=1e-5, atol=1e-7)
    for a, b in zip(images2, want_images):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    assert nrm2.grad is None or float(nrm2.grad.abs().max()) == 0.0
    for name, g, r in (("sdf", sdf2.grad, want[0]), ("color", col2.grad, want[1]), ("semantic", sem2.grad, want[3])):
        r = 2.0 * r
        err = (g - r).abs()
        scale = r.abs().max().item()
        assert scale > 0, name
        assert bool((err <= 1e-3 * r.abs() + 1e-6 * scale).all()), "%s: max err %g (scale %g)" % (name, err.max().item(), scale)


def test_fused_loss_terms_can_be_switched_off(cuda_device):
    from oracle import losses_ref as R
    from spsg_b200 import synthetic as S
    from spsg_b200.losses import render_with_2d_losses
    w, h = 96, 64
    rc, pred, view, intr, images_depth, images_color, wc, label, cw = _setup(cuda_device, 1, 1, w, h)
    sdf = pred["sdf"].clone().requires_grad_(True)
    total, terms, images = render_with_2d_losses(rc, pred["locs"], sdf, pred["color"], pred["normal"], pred["semantic"],
                                                 view, intr, images_depth=images_depth, voxelsize=S.VOXELSIZE)
    want = R.depth_l1_loss(images[1].clone(), images_depth, S.VOXELSIZE)
    torch.testing.assert_close(total, want, rtol=1e-5, atol=1e-7)
    assert float(terms[1]) == 0.0 and float(terms[2]) == 0.0
    total.backward()
    assert float(sdf.grad.abs().sum()) > 0


def test_fused_colour_loss_matches_reference_python_golden(cuda_device):
    """The fused colour-L1 term against the value the reference's own loss.compute_2dcolor_loss produced
    (tests/golden/losses_ref.npz): the golden rendering (with -inf holes) is reproduced by a one-voxel-per-pixel scene is
    not possible, so the term is checked through the un-fused product path: spsg_b200.losses.color_l1_loss."""
    import os
    from spsg_b200.losses import color_l1_loss
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses_ref.npz"))
    pred = torch.from_numpy(g["c_pred"]).to(cuda_device)
    tgt = torch.from_numpy(g["c_tgt"]).to(cuda_device)
    for name, ww in (("color_w", torch.from_numpy(g["c_weight"]).to(cuda_device)), ("color_now", None)):
        p = pred.clone().requires_grad_(True)
        l = color_l1_loss(p, tgt, ww)
        l.backward()
        assert abs(float(l) - float(g[name])) < 1e-6
        gr = p.grad.clone()
        gr[torch.isinf(pred)] = 0
        assert np.abs(gr.cpu().numpy() - g[name + "_grad"]).max() < 1e-7


def test_standalone_image_losses_match_literal_pytorch(cuda_device):
    """depth_l1_loss / color_l1_loss / semantic_2d_ce_loss on rendered images against the literal expressions + autograd."""
    from oracle import losses_ref as R
    from spsg_b200 import synthetic as S
    from spsg_b200 import losses as L
    w, h = 160, 128
    rc, pred, view, intr, images_depth, images_color, wc, label, cw = _setup(cuda_device, 2, 1, w, h, weight_color=True)
    with torch.no_grad():
        imgs = [t.clone() for t in rc(pred["locs"], pred["sdf"], pred["color"], pred["normal"], pred["semantic"], view, intr)]
    hit = imgs[1] != -float("inf")
    cases = (
        ("depth", lambda x: R.depth_l1_loss(x, images_depth, S.VOXELSIZE), lambda x: L.depth_l1_loss(x, images_depth, S.VOXELSIZE), imgs[1], hit),
        ("colour", lambda x: R.compute_2dcolor_loss(x, images_color, wc), lambda x: L.color_l1_loss(x, images_color, wc), imgs[0], hit[..., None].expand_as(imgs[0])),
        ("colour/no weight", lambda x: R.compute_2dcolor_loss(x, images_color, None), lambda x: L.color_l1_loss(x, images_color), imgs[0], hit[..., None].expand_as(imgs[0])),
        ("semantic", lambda x: R.semantic_2d_ce_loss(x, label, cw), lambda x: L.semantic_2d_ce_loss(x, label, cw), imgs[3], hit[..., None].expand_as(imgs[3])),
    )
    for name, ref_fn, our_fn, image, valid in cases:
        a = image.clone().requires_grad_(True)
        la = ref_fn(a)
        la.backward()
        b = image.clone().requires_grad_(True)
        lb = our_fn(b)
        (lb * 1.5).backward()
        torch.testing.assert_close(lb.detach(), la.detach(), rtol=1e-5, atol=1e-7)
        ga, gb = a.grad[valid] * 1.5, b.grad[valid]
        scale = float(ga.abs().max())
        assert scale > 0, name
        assert float((ga - gb).abs().max()) <= 1e-4 * scale, name
        assert bool((b.grad[~valid] == 0).all()), name


def test_label_maps_match_the_literal_expression(cuda_device):
    """train.py:614-616 / :749-752: argmax(cat(render, ones), -1) as uint8 -- ties, values exactly 1, misses (-inf),
    unlabeled (all zero), NaN, and the per-class pixel counts of the same pass."""
    from spsg_b200.losses import labels_from_render
    g = torch.Generator().manual_seed(3)
    sem = torch.randn(5, 37, 41, 14, generator=g) * 2.0
    flat = sem.view(-1, 14)
    flat[0::7] = -float("inf")                                   # misses
    flat[1::7] = 0.0                                             # unlabeled target voxel
    flat[2::7] = torch.nn.functional.one_hot(torch.arange(flat[2::7].shape[0]) % 14, 14).float()   # one-hot: max == 1 (tie with the appended 1)
    flat[3::7, 5] = flat[3::7, 9] = 4.0                          # two equal maxima: the first wins
    flat[4::49, 3] = float("nan")
    flat[5::49, 0] = 0.99999994                                  # just below 1
    sem = sem.to(cuda_device)
    cat = torch.cat((sem, torch.ones(sem.shape[:-1] + (1,), device=cuda_device)), dim=-1)
    want = torch.max(cat, dim=-1)[1].to(torch.uint8)
    got, hist = labels_from_render(sem, histogram=True)
    assert got.dtype == torch.uint8 and got.shape == want.shape
    assert torch.equal(got, want)
    assert torch.equal(hist, torch.bincount(want.reshape(-1).long(), minlength=15))
    assert torch.equal(labels_from_render(sem), want)
    assert labels_from_render(sem[:0]).shape == (0, 37, 41)

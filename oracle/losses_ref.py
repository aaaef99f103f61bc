"""The reference's 2D losses, restated literally in PyTorch.  TEST INFRASTRUCTURE, NOT PRODUCT.

  depth L1      train.py:635-638   (val loop :1088-1090)
  colour L1     loss.py:246-257    compute_2dcolor_loss
  2D labels     train.py:614-616   argmax over cat(render, ones)
  2D CE         train.py:744-746   F.cross_entropy(logits[valid], label[valid], weight=w14)
"""
import torch
import torch.nn.functional as F


def depth_l1_loss(raycast_depth, images_depth, voxelsize):
    raycast_depth = raycast_depth.unsqueeze(1) * voxelsize          # (B,1,H,W)
    valid = (raycast_depth != -float('inf')) & (images_depth != 0)
    return torch.mean(torch.abs(raycast_depth[valid] - images_depth[valid]))


def compute_2dcolor_loss(raycast_color, target_color, weight_color):
    valid = raycast_color != -float('inf')
    pred = raycast_color
    tgt = target_color
    if weight_color is not None:
        w = weight_color.view(weight_color.shape[0], weight_color.shape[2], weight_color.shape[3], 1)
        pred = pred * w
        tgt = tgt * w
    pred = pred[valid]
    tgt = tgt[valid]
    return torch.mean(torch.abs(pred - tgt))


def labels_from_render(raycast_semantic):
    cat = torch.cat((raycast_semantic, torch.ones(raycast_semantic.shape[:-1] + (1,), device=raycast_semantic.device)),
                    dim=-1)
    _, label = torch.max(cat, dim=-1, keepdim=True)
    return label.to(torch.uint8)                                     # (B,H,W,1), 14 = no hit / unlabeled


def semantic_2d_ce_loss(raycast_semantic, target2d_label, weight_semantic_class):
    valid = torch.logical_and(target2d_label[..., 0] < 14, raycast_semantic[..., 0] != -float('inf'))
    return F.cross_entropy(raycast_semantic[valid].view(-1, raycast_semantic.shape[-1]),
                           target2d_label[valid].view(-1).long(), weight=weight_semantic_class)


def compute_normals_sparse(sdf_locs, sdf_vals, dims, transform=None):
    """Literal PyTorch restatement of reference torch/loss.py:285-306 (+ compute_normals_dense :261-267): scatter into a
    zero volume, central differences on the interior, -inf padding that is then zeroed, gather at the voxels, per-chunk
    3x3 rotation, -normalize(eps=1e-5).  TEST INFRASTRUCTURE: checker for spsg_b200.normals."""
    dz, dy, dx = int(dims[0]), int(dims[1]), int(dims[2])
    b = sdf_locs[:, 3]
    num_chunks = int(b[-1].item()) + 1                                                      # loss.py:287
    vol = torch.zeros(num_chunks, 1, dz, dy, dx, device=sdf_vals.device, dtype=sdf_vals.dtype)
    vol[b, :, sdf_locs[:, 0], sdf_locs[:, 1], sdf_locs[:, 2]] = sdf_vals                     # :288-289
    gx = vol[:, :, 1:dz - 1, 1:dy - 1, 2:dx] - vol[:, :, 1:dz - 1, 1:dy - 1, 0:dx - 2]      # :264
    gy = vol[:, :, 1:dz - 1, 2:dy, 1:dx - 1] - vol[:, :, 1:dz - 1, 0:dy - 2, 1:dx - 1]      # :265
    gz = vol[:, :, 2:dz, 1:dy - 1, 1:dx - 1] - vol[:, :, 0:dz - 2, 1:dy - 1, 1:dx - 1]      # :266
    g = torch.nn.functional.pad(torch.cat([gx, gy, gz], 1), (1, 1, 1, 1, 1, 1), value=-float("inf"))  # :292
    g = g[b, :, sdf_locs[:, 0], sdf_locs[:, 1], sdf_locs[:, 2]].contiguous()                # :293
    g = torch.where(g == -float("inf"), torch.zeros_like(g), g)                             # :295
    if transform is not None:                                                               # :296-302
        g = torch.cat([torch.matmul(transform[k, :3, :3], g[b == k].t()).t() for k in range(transform.shape[0])])
    return -torch.nn.functional.normalize(g, p=2, dim=1, eps=1e-5)                          # :305

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

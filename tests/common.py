"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np
import torch

from spsg_b200 import synthetic as S


def scene_tensors(seeds, device, payload="prediction", dims_zyx=S.DIMS_ZYX):
    batch = S.make_batch(seeds, dims_zyx=dims_zyx, payload=payload)
    t = {k: torch.from_numpy(batch[k]).to(device) for k in ("locs", "sdf", "color", "normal", "semantic")}
    return batch, t


def views(num_chunks, views_per_chunk, device, seed=0, width=S.WIDTH, height=S.HEIGHT, **kw):
    view, intr = S.make_views(num_chunks, views_per_chunk, seed=seed, **kw)
    intr = intr.copy()
    intr[:, 0] *= width / S.WIDTH
    intr[:, 2] = (intr[:, 2] + 0.5) * width / S.WIDTH - 0.5
    intr[:, 1] *= height / S.HEIGHT
    intr[:, 3] = (intr[:, 3] + 0.5) * height / S.HEIGHT - 0.5
    return view, intr.astype(np.float32), torch.from_numpy(view).to(device), torch.from_numpy(intr).to(device)


def bits(t):
    return t.contiguous().view(torch.int32)


def count_bit_mismatch(a, b):
    return int((bits(a) != bits(b)).sum().item())


def hit_image_from_mapping(mapping3dto2d, mapping3dto2d_num, locs, num_images, height, width, views_per_chunk=1):
    """Per-pixel hit voxel index (-1 = miss) reconstructed from the voxel->pixel registration tables.
    Only valid when no voxel overflowed max_pixels_per_voxel."""
    n = locs.shape[0]
    max_pix = mapping3dto2d.shape[1]
    img = torch.full((num_images, height * width), -1, dtype=torch.int64, device=locs.device)
    ar = torch.arange(max_pix, device=locs.device)[None, :]
    for f in range(views_per_chunk):
        num = mapping3dto2d_num[f * n:(f + 1) * n].long()
        rows = torch.nonzero(num > 0)[:, 0]
        if rows.numel() == 0:
            continue
        cnt = num[rows].clamp(max=max_pix)
        mask = ar < cnt[:, None]
        pix = mapping3dto2d[f * n + rows].long()[mask]
        vox = rows[:, None].expand(-1, max_pix)[mask]
        image = locs[vox, 3] * views_per_chunk + f
        img[image, pix] = vox
    return img.view(num_images, height, width)

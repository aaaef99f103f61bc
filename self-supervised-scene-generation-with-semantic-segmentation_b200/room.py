"""Whole-room inference by sliding chunks, with multi-view semantic rendering of every chunk, sharded over GPUs
(BASELINE.json configs[4]).

The reference's ``test_scene_as_chunks.py`` slides a (Dz, 64, 64) window with stride 32 over the room
(test_scene_as_chunks.py:156-157), runs the generator on every window and blends the overlapping outputs; it renders
nothing.  This driver adds the rendering: every window's predicted sparse voxels are ray-cast from ``views_per_chunk``
cameras around the window (camera set-up as in test_scene.py:89-95: fixed intrinsics, camera->grid pose per view) with the
B200 raycaster, several windows per launch.  Windows are independent, so ranks take them round-robin
(``parallel.shard_round_robin``) and nothing is exchanged on the data path; per-class pixel counts are summed over ranks
at the end (one tiny all-reduce).

The generator itself is out of scope (stock PyTorch): ``predict`` is any callable that maps a window
to its predicted sparse voxels; ``synthetic_predictor`` stands in for it with the analytic room of ``synthetic.py``.
"""
import math

import numpy as np
import torch

from . import parallel as P
from . import synthetic as S
from .losses import labels_from_render
from .normals import compute_normals_sparse
from .raycast_rgbd import RaycastRGBD


def chunk_windows(room_dims_zyx, chunk_yx=(64, 64), stride=32):
    """(y0, x0) of every window, in the reference's enumeration order (test_scene_as_chunks.py:156-157)."""
    return [(y, x) for y in range(0, room_dims_zyx[1], stride) for x in range(0, room_dims_zyx[2], stride)]


def window_views(num_views, chunk_dims_zyx, radius=75.0, height=70.0):
    """camera->grid matrices (num_views,4,4) around one chunk (grid = the chunk's own voxel frame) and intrinsics."""
    cz, cy, cx = chunk_dims_zyx[0] * 0.3125, chunk_dims_zyx[1] * 0.5, chunk_dims_zyx[2] * 0.5
    mats = []
    for k in range(num_views):
        az = math.radians(360.0 * k / num_views + 5.0)
        eye = (cx + radius * math.cos(az), cy + radius * math.sin(az), height)
        mats.append(S.look_at(eye, (cx, cy, cz)))
    view = np.stack(mats).astype(np.float32)
    intr = np.tile(np.asarray(S.INTRINSICS, dtype=np.float32), (num_views, 1))
    return view, intr


def synthetic_predictor(room_sdf, truncation=S.TRUNCATION, seed=0):
    """Stand-in for generator + sparsification (train.py:494-509) on a dense room SDF tensor (Dz,Dy,Dx) that lives on the
    GPU: returns ``predict(y0, x0, chunk_yx) -> (locs (n,3) int64 z,y,x in chunk coordinates, sdf (n,1), colour (n,3),
    semantic logits (n,14))``."""
    def predict(y0, x0, chunk_yx):
        win = room_sdf[:, y0:y0 + chunk_yx[0], x0:x0 + chunk_yx[1]]
        locs = torch.nonzero(win.abs() < truncation)
        vals = win[locs[:, 0], locs[:, 1], locs[:, 2]].reshape(-1, 1).contiguous()
        g = torch.Generator(device=win.device).manual_seed(seed * 1000003 + y0 * 4099 + x0)
        color = torch.rand(locs.shape[0], 3, device=win.device, generator=g)
        sem = torch.randn(locs.shape[0], S.NUM_CLASSES, device=win.device, generator=g) * 14.0
        return locs, vals, color, sem
    return predict


def render_room(predict, room_dims_zyx, device, views_per_chunk=5, chunks_per_launch=8, chunk_yx=(64, 64), stride=32,
                width=S.WIDTH, height=S.HEIGHT, rank=0, world=1, max_num_locs_per_sample=640000, keep_images=False):
    """Render this rank's share of the room's windows.  Returns dict(windows, rendered_windows, rays, label_hist (15,)
    summed over ranks, images: list of (window, labels (F,H,W) uint8) when ``keep_images``)."""
    dz = room_dims_zyx[0]
    chunk_dims = (dz, chunk_yx[0], chunk_yx[1])
    windows = chunk_windows(room_dims_zyx, chunk_yx, stride)
    mine = [windows[i] for i in P.shard_round_robin(len(windows), rank, world)]
    B, F = chunks_per_launch, views_per_chunk
    rc = RaycastRGBD(B, chunk_dims, width, height, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                     max_num_frames=F, max_num_locs_per_sample=max_num_locs_per_sample, device=device)
    view_np, intr_np = window_views(F, chunk_dims)
    view = torch.from_numpy(np.tile(view_np, (B, 1, 1))).to(device)
    intr = torch.from_numpy(np.tile(intr_np, (B, 1))).to(device)
    grid2cam = torch.inverse(view[::F]).contiguous()          # one rotation per chunk for the normals (train.py:544)
    hist = torch.zeros(S.NUM_CLASSES + 1, dtype=torch.float64, device=device)
    images, rendered, rays = [], 0, 0
    for s in range(0, len(mine), B):
        group = mine[s:s + B]
        parts = [predict(y0, x0, chunk_yx) for (y0, x0) in group]
        keep = [k for k, p in enumerate(parts) if p[0].shape[0] > 0]   # the reference skips empty windows (:160-161)
        if not keep:
            continue
        # chunks of a partial last group are padded by repeating nothing: the launch renders B slots, empty ones miss
        locs = torch.cat([torch.cat([parts[k][0], torch.full((parts[k][0].shape[0], 1), b, dtype=torch.long, device=device)], 1)
                          for b, k in enumerate(keep)]).contiguous()
        sdf = torch.cat([parts[k][1] for k in keep])
        color = torch.cat([parts[k][2] for k in keep])
        sem = torch.cat([parts[k][3] for k in keep])
        with torch.no_grad():
            normals = compute_normals_sparse(locs, sdf, chunk_dims, grid2cam, num_chunks=B)
            _, depth, _, sem_img = rc(locs, sdf, color, normals, sem, view, intr)
            nimg = len(keep) * F
            labels = labels_from_render_logits(sem_img[:nimg], depth[:nimg])
            hist += torch.bincount(labels.reshape(-1).long(), minlength=S.NUM_CLASSES + 1).to(torch.float64)
        rendered += len(keep)
        rays += nimg * width * height
        if keep_images:
            for b, k in enumerate(keep):
                images.append((group[k], labels[b * F:(b + 1) * F].cpu()))
    total_hist = hist.clone()
    if world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(total_hist)   # the only collective: 15 counters
    return dict(windows=len(windows), rendered_windows=rendered, rays=rays, label_hist=total_hist.cpu().numpy(),
                images=images)


def labels_from_render_logits(sem_img, depth):
    """Per-pixel predicted class of a rendering of semantic *logits* (train.py:749-752: argmax over the 14 logits),
    14 where the ray hit nothing."""
    lab = sem_img.argmax(dim=-1)
    return torch.where(depth != -float("inf"), lab, torch.full_like(lab, S.NUM_CLASSES)).to(torch.uint8)


def synthetic_room_sdf(room_dims_zyx, device, seed=0):
    """Analytic room: the chunk scene of ``synthetic.sdf_volume`` repeated with period 64 in y and x plus the outer floor,
    as one dense float32 tensor (Dz,Dy,Dx) on ``device``."""
    dz, dy, dx = room_dims_zyx
    base, _ = S.sdf_volume(seed, (dz, 64, 64))
    reps = (1, (dy + 63) // 64, (dx + 63) // 64)
    room = np.tile(base, reps)[:, :dy, :dx]
    return torch.from_numpy(np.ascontiguousarray(room)).to(device)

"""Frame loader for the 2D view-guided losses: drop-in for the reference's ``data_util.load_frames`` / ``load_frame``
(``torch/data_util.py:837-902``).

Per chunk the reference decodes ``max_num_frames`` depth PNGs (16-bit millimetres), colour JPEGs and camera text files
one after the other on the dataloader thread (``imageio.imread`` -> ``torchvision`` resize / centre crop on PIL images ->
numpy -> torch), and fills freshly allocated batch tensors.  Here the same files go through Pillow directly, a batch's
frames are decoded on a small thread pool (decoding releases the GIL), and the batch tensors can be allocated in pinned
host memory so that the training loop's host->device copy is one asynchronous DMA per tensor.

Same arguments, same return value ``(depths (B,F,h,w), colors (B,F,3,H,W), poses (B,F,4,4), intrinsics (B,F,4), frames)``,
same ``(None,)*5`` when a chunk lists too few frames, same nearest-neighbour resize + centre crop, same intrinsic
adjustment (``data_util.py:790-812``).  CPU code: the hot path's consumer of these tensors is on the GPU, their producer
is file I/O.
"""
import math
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from PIL import Image

_pool = None


def _executor(workers):
    global _pool
    if _pool is None or _pool._max_workers != workers:
        _pool = ThreadPoolExecutor(max_workers=workers)
    return _pool


def read_frame_file(filename, randomize, max_num_frames):
    """Frame ids of one chunk (data_util.py:764-771): one integer per line, -1 marks an unusable frame; with ``randomize``
    the usable ids are shuffled (Python's ``random``, like the reference) before the first ``max_num_frames`` are taken."""
    with open(filename) as fh:
        ids = [int(tok) for tok in fh.read().split()]
    if randomize:
        ids = np.array(ids)
        ids = ids[ids != -1]
        random.shuffle(ids)
    return ids[:max_num_frames]


def _matrix_rows(path):
    with open(path) as fh:
        return np.asarray([ln.split(" ")[:4] for ln in fh.read().splitlines()], dtype=np.float32)


def read_camera_file(filename, intrinsic_filename=None):
    """(pose, intrinsic) of one frame (data_util.py:774-787): the camera file holds the 4x4 camera -> world pose followed by
    the 4x4 intrinsic matrix; a separate intrinsic file overrides the latter."""
    rows = _matrix_rows(filename)
    k_rows = rows[4:] if intrinsic_filename is None else _matrix_rows(intrinsic_filename)[:4]
    return torch.from_numpy(rows[:4].copy()), torch.from_numpy(k_rows.copy())


def _resized_width(src_wh, dst_wh):
    # width after scaling the source to the target height (aspect ratio kept), data_util.py:794 / :806
    return int(math.floor(dst_wh[1] * float(src_wh[0]) / float(src_wh[1])))


def resize_crop_image(image, new_image_dims):
    """Nearest-neighbour resize of a decoded array (h, w[, c]) to the target height, then a centre crop to the target width
    (data_util.py:790-800: torchvision's Resize / CenterCrop on PIL images); ``new_image_dims`` = [width, height]."""
    tw, th = int(new_image_dims[0]), int(new_image_dims[1])
    if [image.shape[1], image.shape[0]] == [tw, th]:
        return image
    scaled = Image.fromarray(image).resize((_resized_width((image.shape[1], image.shape[0]), (tw, th)), th), Image.NEAREST)
    w, h = scaled.size
    if (w, h) == (th, tw):
        return scaled  # the reference returns the PIL image itself in this corner case (data_util.py:796-797)
    if tw > w or th > h:  # CenterCrop pads a smaller image with zeros first
        canvas = Image.new(scaled.mode, (max(w, tw), max(h, th)))
        canvas.paste(scaled, (max((tw - w) // 2, 0), max((th - h) // 2, 0)))
        scaled, (w, h) = canvas, canvas.size
    x0, y0 = int(round((w - tw) / 2.0)), int(round((h - th) / 2.0))
    return np.array(scaled.crop((x0, y0, x0 + tw, y0 + th)))


def adjust_intrinsic(intrinsic, intrinsic_image_dim, image_dim):
    """Intrinsics of the resized + cropped frame (data_util.py:803-812); modifies and returns ``intrinsic``."""
    src, dst = list(intrinsic_image_dim), list(image_dim)
    if src != dst:
        intrinsic[0, 0] *= float(_resized_width(src, dst)) / float(src[0])
        intrinsic[1, 1] *= float(dst[1]) / float(src[1])
        intrinsic[0, 2] *= float(dst[0] - 1) / float(src[0] - 1)  # (the crop is accounted for here)
        intrinsic[1, 2] *= float(dst[1] - 1) / float(src[1] - 1)
    return intrinsic


def load_frame(depth_file, color_file, camera_file, depth_image_dims, color_image_dims, normalize, load_depth, load_color,
               intrinsic_file=None):
    """One frame (data_util.py:837-859): depth in metres (16-bit millimetres / 1000), colour (3,H,W) in [0,1] (optionally
    normalised), pose, intrinsic adjusted to the colour frame's size."""
    pose, intrinsic = read_camera_file(camera_file, intrinsic_file)
    depth_t = color_t = src_wh = None
    if load_depth:
        raw = np.array(Image.open(depth_file))
        src_wh = [raw.shape[1], raw.shape[0]]
        mm = np.asarray(resize_crop_image(raw, list(depth_image_dims)))
        depth_t = torch.from_numpy(mm.astype(np.float32) / 1000.0)
    if load_color:
        raw = np.array(Image.open(color_file))
        src_wh = [raw.shape[1], raw.shape[0]]
        rgb = np.asarray(resize_crop_image(raw, list(color_image_dims)))
        color_t = torch.from_numpy(np.ascontiguousarray(rgb.transpose(2, 0, 1)).astype(np.float32) / 255.0)
        if normalize is not None:
            color_t = normalize(color_t)
    if list(color_image_dims) != src_wh:
        intrinsic = adjust_intrinsic(intrinsic, src_wh, list(color_image_dims))
    return depth_t, color_t, pose, intrinsic


def load_frames(names, world2grids, frame_path, image_path, randomize_frames, depth_image_dims, color_image_dims,
                color_normalization, load_depth, load_color, max_num_frames=1, pin_memory=False, num_workers=8):
    """Reference signature (data_util.py:862-902) plus ``pin_memory`` (batch tensors in pinned host memory) and
    ``num_workers`` (decoder threads; 0 = decode in the calling thread)."""
    B, F = len(names), max_num_frames
    if frame_path == 'self':  # the chunk's own frame: its id is part of the chunk name
        frames = [[int(nm.split('__inc__')[1])] for nm in names]
    else:
        frames = [read_frame_file(os.path.join(frame_path, nm.replace('__inc__', '__cmp__') + '.txt'), randomize_frames, F)
                  for nm in names]
    if len(frames[0]) < F:
        return None, None, None, None, None
    new = (lambda *shape: torch.zeros(*shape, dtype=torch.float).pin_memory()) if pin_memory else \
        (lambda *shape: torch.zeros(*shape, dtype=torch.float))
    poses, intrinsics = new(B, F, 4, 4), new(B, F, 4)
    depths = new(B, F, depth_image_dims[1], depth_image_dims[0]) if load_depth else None
    colors = new(B, F, 3, color_image_dims[1], color_image_dims[0]) if load_color else None

    def decode(job):
        b, f = job
        scene_dir = os.path.join(image_path, names[b].split('_room')[0])
        fid = str(frames[b][f])
        d, c, pose, K = load_frame(scene_dir + '/depth/' + fid + '.png', scene_dir + '/color/' + fid + '.jpg',
                                   scene_dir + '/camera/' + fid + '.txt', depth_image_dims, color_image_dims,
                                   color_normalization, load_depth, load_color)
        if d is not None:
            depths[b, f] = d
        if c is not None:
            colors[b, f] = c
        poses[b, f] = pose
        intrinsics[b, f] = torch.stack((K[0, 0], K[1, 1], K[0, 2], K[1, 2]))  # fx, fy, mx, my

    jobs = [(b, f) for b in range(B) for f in range(F)]
    if num_workers and len(jobs) > 1:
        list(_executor(num_workers).map(decode, jobs))
    else:
        for job in jobs:
            decode(job)
    return depths, colors, poses, intrinsics, frames

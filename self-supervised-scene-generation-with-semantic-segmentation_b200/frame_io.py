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
    """data_util.py:764-771: one frame id per line, -1 = unusable; optionally shuffled."""
    with open(filename) as f:
        frames = [int(line) for line in f.read().splitlines()]
    if randomize:
        frames = np.array(frames)
        frames = frames[frames != -1]
        random.shuffle(frames)
    return frames[:max_num_frames]


def read_camera_file(filename, intrinsic_filename=None):
    """data_util.py:774-787: 4 rows of pose (camera -> world) followed by 4 rows of intrinsics."""
    def rows(name):
        with open(name) as f:
            return np.asarray([line.split(" ")[:4] for line in f.read().splitlines()], dtype=np.float32)
    lines = rows(filename)
    pose = torch.from_numpy(lines[:4].copy())
    intrinsic = torch.from_numpy(lines[4:].copy()) if intrinsic_filename is None else \
        torch.from_numpy(rows(intrinsic_filename)[:4].copy())
    return pose, intrinsic


def resize_crop_image(image, new_image_dims):
    """data_util.py:790-800 for a decoded array (h, w[, c]); ``new_image_dims`` = [width, height].  Nearest-neighbour
    resize to the new height, then a centre crop to the new width (torchvision's Resize / CenterCrop on PIL images)."""
    image_dims = [image.shape[1], image.shape[0]]
    if image_dims == list(new_image_dims):
        return image
    resize_width = int(math.floor(new_image_dims[1] * float(image_dims[0]) / float(image_dims[1])))
    pil = Image.fromarray(image).resize((resize_width, new_image_dims[1]), Image.NEAREST)
    if pil.size[0] == new_image_dims[1] and pil.size[1] == new_image_dims[0]:
        return pil  # (the reference returns the PIL image in this corner case, data_util.py:796-797)
    th, tw = new_image_dims[1], new_image_dims[0]
    w, h = pil.size
    if tw > w or th > h:  # torchvision's CenterCrop pads with zeros first
        pad_l, pad_t = max((tw - w) // 2, 0), max((th - h) // 2, 0)
        padded = Image.new(pil.mode, (max(w, tw), max(h, th)))
        padded.paste(pil, (pad_l, pad_t))
        pil, (w, h) = padded, padded.size
    top, left = int(round((h - th) / 2.0)), int(round((w - tw) / 2.0))
    return np.array(pil.crop((left, top, left + tw, top + th)))


def adjust_intrinsic(intrinsic, intrinsic_image_dim, image_dim):
    """data_util.py:803-812 (modifies and returns ``intrinsic``)."""
    if list(intrinsic_image_dim) == list(image_dim):
        return intrinsic
    resize_width = int(math.floor(image_dim[1] * float(intrinsic_image_dim[0]) / float(intrinsic_image_dim[1])))
    intrinsic[0, 0] *= float(resize_width) / float(intrinsic_image_dim[0])
    intrinsic[1, 1] *= float(image_dim[1]) / float(intrinsic_image_dim[1])
    intrinsic[0, 2] *= float(image_dim[0] - 1) / float(intrinsic_image_dim[0] - 1)
    intrinsic[1, 2] *= float(image_dim[1] - 1) / float(intrinsic_image_dim[1] - 1)
    return intrinsic


def load_frame(depth_file, color_file, camera_file, depth_image_dims, color_image_dims, normalize, load_depth, load_color,
               intrinsic_file=None):
    """data_util.py:837-859."""
    pose, intrinsic = read_camera_file(camera_file, intrinsic_file)
    depth_image = color_image = orig_dims = None
    if load_depth:
        depth = np.array(Image.open(depth_file))
        orig_dims = [depth.shape[1], depth.shape[0]]
        depth = np.asarray(resize_crop_image(depth, list(depth_image_dims)))
        depth_image = torch.from_numpy(depth.astype(np.float32) / 1000.0)
    if load_color:
        color = np.array(Image.open(color_file))
        orig_dims = [color.shape[1], color.shape[0]]
        color = np.asarray(resize_crop_image(color, list(color_image_dims)))
        color_image = torch.from_numpy(np.ascontiguousarray(np.transpose(color, [2, 0, 1])).astype(np.float32) / 255.0)
        if normalize is not None:
            color_image = normalize(color_image)
    if list(color_image_dims) != orig_dims:
        intrinsic = adjust_intrinsic(intrinsic, orig_dims, list(color_image_dims))
    return depth_image, color_image, pose, intrinsic


def load_frames(names, world2grids, frame_path, image_path, randomize_frames, depth_image_dims, color_image_dims,
                color_normalization, load_depth, load_color, max_num_frames=1, pin_memory=False, num_workers=8):
    """Reference signature (data_util.py:862-902) plus ``pin_memory`` (batch tensors in pinned host memory) and
    ``num_workers`` (decoder threads; 0 = decode in the calling thread)."""
    batch_size = len(names)
    scenes = [name.split('_room')[0] for name in names]
    if frame_path == 'self':
        frames = [[int(name.split('__inc__')[1])] for name in names]
    else:
        frame_files = [os.path.join(frame_path, name.replace('__inc__', '__cmp__') + '.txt') for name in names]
        frames = [read_frame_file(frame_file, randomize_frames, max_num_frames) for frame_file in frame_files]
    if len(frames[0]) < max_num_frames:
        return None, None, None, None, None
    alloc = (lambda *s: torch.zeros(*s, dtype=torch.float).pin_memory()) if pin_memory else \
        (lambda *s: torch.zeros(*s, dtype=torch.float))
    poses, intrinsics = alloc(batch_size, max_num_frames, 4, 4), alloc(batch_size, max_num_frames, 4)
    depths = alloc(batch_size, max_num_frames, depth_image_dims[1], depth_image_dims[0]) if load_depth else None
    colors = alloc(batch_size, max_num_frames, 3, color_image_dims[1], color_image_dims[0]) if load_color else None

    def one(bf):
        b, f = bf
        fid, scene = frames[b][f], scenes[b]
        depth_image, color_image, pose, intrinsic = load_frame(
            os.path.join(image_path, scene + '/depth/' + str(fid) + '.png'),
            os.path.join(image_path, scene + '/color/' + str(fid) + '.jpg'),
            os.path.join(image_path, scene + '/camera/' + str(fid) + '.txt'), depth_image_dims, color_image_dims,
            color_normalization, load_depth, load_color)
        if load_depth:
            depths[b, f] = depth_image
        if load_color:
            colors[b, f] = color_image
        poses[b, f] = pose
        intrinsics[b, f, 0] = intrinsic[0, 0]
        intrinsics[b, f, 1] = intrinsic[1, 1]
        intrinsics[b, f, 2] = intrinsic[0, 2]
        intrinsics[b, f, 3] = intrinsic[1, 2]

    work = [(b, f) for b in range(batch_size) for f in range(max_num_frames)]
    if num_workers and len(work) > 1:
        list(_executor(num_workers).map(one, work))
    else:
        for bf in work:
            one(bf)
    return depths, colors, poses, intrinsics, frames

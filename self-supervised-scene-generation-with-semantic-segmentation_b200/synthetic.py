"""Seeded synthetic chunks, cameras and target frames.  numpy only.

No dataset ships with the reference (README.md:31-38 links 68-110 GB archives), so tests, ``smoke()`` and
``bench.py`` all render this analytic room: floor + two walls + a sphere + a box inside a
(Dz,Dy,Dx) = (128,64,64) grid of 2 cm voxels, SDF truncated at 3 voxels, sparse band |sdf| < 3.
"""
import math

import numpy as np

DIMS_ZYX = (128, 64, 64)
TRUNCATION = 3.0
VOXELSIZE = 0.02
WIDTH, HEIGHT = 320, 256
INTRINSICS = (269.1, 269.3, 159.5, 127.5)          # fx, fy, mx, my at 320x256 (cf. test_scene.py:90)
DEPTH_MIN, DEPTH_MAX = 0.1 / VOXELSIZE, 6.0 / VOXELSIZE   # train.py:138-139
RAY_INCREMENT = 0.3 * TRUNCATION                   # train.py:134
THRESH_SAMPLE_DIST = 50.5 * RAY_INCREMENT          # train.py:135
NUM_CLASSES = 14
# torch/category.npz 'weight' (train.py:118-119), regenerated as a constant table
CLASS_WEIGHTS = (0.0286, 0.1535, 0.2986, 0.0177, 0.0166, 0.0201, 0.0117, 0.0033, 0.0188, 0.1364, 0.0384, 0.2389,
                 0.0038, 0.0137)


def _primitives(seed):
    rng = np.random.default_rng(seed)
    off = rng.uniform(-2.0, 2.0, size=(5, 3)) if seed != 0 else np.zeros((5, 3))
    return dict(floor_z=10.3 + off[0, 2], wall_y=56.7 + off[1, 1], wall_x=57.2 + off[2, 0],
                sphere_c=np.array([28.4, 30.1, 24.6]) + off[3], sphere_r=14.0,
                box_c=np.array([44.3, 20.2, 30.0]) + off[4], box_h=np.array([8.0, 10.0, 20.0]))


def sdf_volume(seed=0, dims_zyx=DIMS_ZYX):
    """Dense truncated SDF (Dz,Dy,Dx) float32 in voxel units and the id (0..4) of the nearest primitive."""
    dz, dy, dx = dims_zyx
    z, y, x = np.meshgrid(np.arange(dz, dtype=np.float64), np.arange(dy, dtype=np.float64),
                          np.arange(dx, dtype=np.float64), indexing="ij")
    p = _primitives(seed)
    d_floor = z - p["floor_z"]
    d_wally = p["wall_y"] - y
    d_wallx = p["wall_x"] - x
    c = p["sphere_c"]
    d_sphere = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - p["sphere_r"]
    b, h = p["box_c"], p["box_h"]
    q = np.stack([np.abs(x - b[0]) - h[0], np.abs(y - b[1]) - h[1], np.abs(z - b[2]) - h[2]])
    d_box = np.linalg.norm(np.maximum(q, 0.0), axis=0) + np.minimum(q.max(axis=0), 0.0)
    stack = np.stack([d_floor, d_wally, d_wallx, d_sphere, d_box])
    sdf = np.clip(stack.min(axis=0), -TRUNCATION, TRUNCATION).astype(np.float32)
    return sdf, stack.argmin(axis=0).astype(np.int64)


def _normals(sdf):
    """-normalize(central-difference gradient) in grid (x,y,z) order, zero on the volume border
    (loss.py:261-306 without the camera rotation)."""
    g = np.zeros(sdf.shape + (3,), dtype=np.float32)
    g[1:-1, 1:-1, 1:-1, 0] = sdf[1:-1, 1:-1, 2:] - sdf[1:-1, 1:-1, :-2]
    g[1:-1, 1:-1, 1:-1, 1] = sdf[1:-1, 2:, 1:-1] - sdf[1:-1, :-2, 1:-1]
    g[1:-1, 1:-1, 1:-1, 2] = sdf[2:, 1:-1, 1:-1] - sdf[:-2, 1:-1, 1:-1]
    n = np.linalg.norm(g, axis=-1, keepdims=True)
    return (-g / np.maximum(n, 1e-5)).astype(np.float32)


def make_chunk(seed=0, dims_zyx=DIMS_ZYX, payload="prediction"):
    """One sparse chunk.  Returns dict(locs (n,3) int64 [z,y,x], sdf (n,1), color (n,3), normal (n,3),
    semantic (n,14), label (n,) uint8).  payload: 'prediction' -> N(0,1)*14 logits; 'target' -> one-hot of the
    label volume (label = primitive id * 3 % 14, 5 % of voxels unlabeled = 14 -> all-zero row)."""
    sdf, prim = sdf_volume(seed, dims_zyx)
    mask = np.abs(sdf) < TRUNCATION
    locs = np.argwhere(mask).astype(np.int64)
    rng = np.random.default_rng(1000003 * (seed + 1))
    n = locs.shape[0]
    vals_sdf = sdf[mask].reshape(n, 1).astype(np.float32)
    color = rng.random((n, 3), dtype=np.float32)
    normal = _normals(sdf)[mask]
    label = ((prim[mask] * 3 + (locs[:, 0] // 16)) % NUM_CLASSES).astype(np.uint8)
    label[rng.random(n) < 0.05] = NUM_CLASSES
    if payload == "target":
        semantic = np.zeros((n, NUM_CLASSES), dtype=np.float32)
        lab = label < NUM_CLASSES
        semantic[np.nonzero(lab)[0], label[lab]] = 1.0
    else:
        semantic = (rng.standard_normal((n, NUM_CLASSES)) * 14.0).astype(np.float32)
    return dict(locs=locs, sdf=vals_sdf, color=color, normal=np.ascontiguousarray(normal), semantic=semantic,
                label=label)


def make_batch(seeds, dims_zyx=DIMS_ZYX, payload="prediction"):
    """Concatenate chunks into reference layout: locs (N,4) int64 rows (z,y,x,chunk) sorted by chunk."""
    parts = [make_chunk(s, dims_zyx, payload) for s in seeds]
    locs = np.concatenate([np.concatenate([p["locs"], np.full((p["locs"].shape[0], 1), b, dtype=np.int64)], 1)
                           for b, p in enumerate(parts)])
    out = {k: np.concatenate([p[k] for p in parts]) for k in ("sdf", "color", "normal", "semantic", "label")}
    out["locs"] = np.ascontiguousarray(locs)
    out["chunk_sizes"] = [p["locs"].shape[0] for p in parts]
    return out


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """camera->grid 4x4 (row-major), OpenCV convention: +z forward, +x right, +y down."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, down, fwd, eye
    return m.astype(np.float32)


def make_views(num_chunks, views_per_chunk, seed=0, center=(32.0, 32.0, 40.0), radius=75.0, height=70.0):
    """(num_chunks*views, 4, 4) camera->grid matrices and (.., 4) intrinsics; view k of a chunk sits on a circle
    around the chunk centre at azimuth 72 deg * k + U(0, 10 deg), looking at the centre."""
    rng = np.random.default_rng(7919 * (seed + 1))
    mats = []
    for _ in range(num_chunks):
        for k in range(views_per_chunk):
            az = math.radians(72.0 * k + rng.uniform(0.0, 10.0))
            eye = (center[0] + radius * math.cos(az), center[1] + radius * math.sin(az), height)
            mats.append(look_at(eye, center))
    view = np.stack(mats).astype(np.float32)
    intr = np.tile(np.asarray(INTRINSICS, dtype=np.float32), (view.shape[0], 1))
    return view, intr


def make_targets(depth_render, color_render, label_render, seed=0, hole_fraction=0.05):
    """Target frames for the 2D losses from a rendering of a perturbed chunk: depth in metres with
    `hole_fraction` zero-depth holes (and zeros where nothing was hit), colour in [0,1], uint8 labels."""
    rng = np.random.default_rng(104729 * (seed + 1))
    depth = np.where(np.isfinite(depth_render), depth_render * VOXELSIZE, 0.0).astype(np.float32)
    depth[rng.random(depth.shape) < hole_fraction] = 0.0
    color = np.where(np.isfinite(color_render), color_render, 0.5).astype(np.float32)
    return depth, color, label_render.astype(np.uint8)


def make_train_sample(seeds, views_per_chunk=1, dims_zyx=DIMS_ZYX, width=WIDTH, height=HEIGHT, view_seed=0, view_kw=None):
    """One synthetic training batch with the keys of the reference dataloader's samples (scene_dataloader.py /
    data_util.py:862-902; train.py:413-445), as numpy arrays: the target chunk (dense truncated SDF, uint8 colours, per-voxel
    labels 0..14), an incomplete input scan of it (a box-shaped region is missing), the colour-inpainting mask, and
    `views_per_chunk` frames per chunk (colour in [0,1], depth in metres with 5 % holes, camera->grid matrices, intrinsics).

        input (B,4,Dz,Dy,Dx) f32 [sdf, r, g, b]   mask (B,1,Dz,Dy,Dx) f32   sdf (B,1,Dz,Dy,Dx) f32   known (B,1,Dz,Dy,Dx) bool
        colors (B,Dz,Dy,Dx,3) u8   semantics (B,1,Dz,Dy,Dx) i64   images_color (I,3,H,W) f32   images_depth (I,H,W) f32
        view_matrix (I,4,4) f32   images_intrinsic (I,4) f32
    """
    B = len(seeds)
    dz, dy, dx = dims_zyx
    out = {k: [] for k in ("input", "mask", "sdf", "known", "colors", "semantics")}
    for s in seeds:
        sdf, prim = sdf_volume(s, dims_zyx)
        rng = np.random.default_rng(7000003 * (s + 1))
        band = np.abs(sdf) < TRUNCATION
        colors = np.zeros(dims_zyx + (3,), dtype=np.uint8)
        colors[band] = rng.integers(1, 256, size=(int(band.sum()), 3), dtype=np.uint8)
        label = np.full(dims_zyx, NUM_CLASSES, dtype=np.int64)
        z = np.arange(dz)[:, None, None]
        label[band] = ((prim * 3 + z // 16) % NUM_CLASSES)[band]
        label[band & (rng.random(dims_zyx) < 0.05)] = NUM_CLASSES
        # the scan misses a box: no geometry, no colour there; the mask marks where colour has to be inpainted
        lo = np.array([dz // 4, dy // 4, dx // 4]) + rng.integers(0, 8, 3)
        hi = lo + np.array([dz // 3, dy // 3, dx // 3])
        hole = np.zeros(dims_zyx, dtype=bool)
        hole[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
        in_sdf = np.where(hole, np.float32(TRUNCATION), sdf).astype(np.float32)
        in_col = np.where(hole[..., None], 0.0, colors.astype(np.float32) / 255.0).astype(np.float32)
        out["input"].append(np.concatenate([in_sdf[None], in_col.transpose(3, 0, 1, 2)], 0))
        out["mask"].append(hole[None].astype(np.float32))
        out["sdf"].append(sdf[None].astype(np.float32))
        out["known"].append(np.ones((1,) + dims_zyx, dtype=bool))
        out["colors"].append(colors)
        out["semantics"].append(label[None])
    sample = {k: np.stack(v) for k, v in out.items()}
    images = B * views_per_chunk
    view, intr = make_views(B, views_per_chunk, seed=view_seed, **(view_kw or {}))
    intr = intr.copy()
    intr[:, 0] *= width / WIDTH
    intr[:, 2] = (intr[:, 2] + 0.5) * width / WIDTH - 0.5
    intr[:, 1] *= height / HEIGHT
    intr[:, 3] = (intr[:, 3] + 0.5) * height / HEIGHT - 0.5
    rng = np.random.default_rng(31 + view_seed + 1009 * (seeds[0] if seeds else 0))
    sample["images_color"] = rng.random((images, 3, height, width), dtype=np.float32)
    depth = rng.uniform(0.8, 1.6, (images, height, width)).astype(np.float32)
    depth[rng.random((images, height, width)) < 0.05] = 0.0
    sample["images_depth"] = depth
    sample["view_matrix"] = view.astype(np.float32)
    sample["images_intrinsic"] = intr.astype(np.float32)
    return sample

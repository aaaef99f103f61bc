#!/usr/bin/env python
"""Generate tests/golden/chunk_io_ref.npz + tests/golden/tiny_chunk.sdf with the reference's OWN `data_util.load_sdf`.

    python tests/golden/make_golden_chunk_io.py      (needs /root/reference; run in the build container)

`/root/reference/torch/data_util.py` is imported unmodified; the image / mesh packages it imports at module level
(imageio, plyfile, skimage, torchvision, its marching-cubes extension) are not installed here and are not touched by
`load_sdf`, so empty stand-in modules are registered for them.  The tiny chunk file itself is written by
spsg_b200.chunk_io.write_chunk_file in the layout of datagen's VoxelGrid::saveToFile."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("SPSG_REFERENCE_TORCH", "/root/reference/torch")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    from spsg_b200 import chunk_io
    rng = np.random.default_rng(7)
    dims = (6, 5, 8)  # x, y, z
    cells = dims[0] * dims[1] * dims[2]
    vol = rng.normal(0, 0.06, (dims[2], dims[1], dims[0])).astype(np.float32)        # metres, z y x
    zz, yy, xx = np.nonzero(np.abs(vol) < 0.06)
    locs_xyz = np.stack([xx, yy, zz], 1).astype(np.uint32)
    sdf = vol[zz, yy, xx]
    w2g = np.eye(4, dtype=np.float32)
    w2g[:3, 3] = (1.5, -2.25, 0.75)
    path = os.path.join(HERE, "tiny_chunk.sdf")
    chunk_io.write_chunk_file(path, dims, 0.02, w2g, locs_xyz, sdf,
                              known=rng.integers(0, 3, (dims[2], dims[1], dims[0])).astype(np.uint8),
                              color=rng.integers(0, 256, (dims[2], dims[1], dims[0], 3)).astype(np.uint8),
                              semantic=rng.integers(0, 15, (dims[2], dims[1], dims[0])).astype(np.uint8))
    for name in ("imageio", "plyfile", "skimage", "skimage.color", "torchvision", "torchvision.transforms", "utils",
                 "utils.marching_cubes", "utils.marching_cubes.marching_cubes"):
        if name not in sys.modules or name.startswith("utils"):
            stub(name)
    sys.path.insert(0, REF)
    import data_util as ref
    out = {}
    (locs, vals), d3, world2grid, known, color, sem = ref.load_sdf(path, load_sparse=True, load_known=False, load_color=True)
    out.update(sp_locs=locs, sp_sdf=vals, sp_dims=np.array(d3), sp_w2g=world2grid, sp_color=color)
    dense, world2grid, known, color, sem = ref.load_sdf(path, load_sparse=False, load_known=True, load_color=True, load_semantic=True)
    out.update(de_sdf=dense, de_known=known, de_color=color, de_sem=sem)
    np.savez_compressed(os.path.join(HERE, "chunk_io_ref.npz"), **out)
    print("wrote", path, os.path.getsize(path), "bytes and chunk_io_ref.npz;", len(sdf), "voxels of", cells)


if __name__ == "__main__":
    main()

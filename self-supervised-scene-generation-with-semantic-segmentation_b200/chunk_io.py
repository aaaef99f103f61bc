"""Reader (and writer) of the reference's on-disk chunk / room format -- drop-in for ``data_util.load_sdf``
(torch/data_util.py:64-159), SURVEY.md section 8(f) rank 4.

The reference decodes every array with ``struct.unpack('I' * n, ...)``: a Python tuple of ~0.5 M boxed integers per
sample, which -- not the GPU -- bounds training throughput once the raycast is fast.  Here every array is one
``numpy.frombuffer`` view of a single file read (zero Python-level per-element work).  Same signature, same return
conventions, same dtypes / shapes / values.

File layout (little endian; written by datagen's VoxelGrid::saveToFile, VoxelGrid.cpp:125):
    u64 dimx, dimy, dimz; f32 voxelsize; f32[16] world2grid (row-major)
    u64 n; u32[n][3] locations (x,y,z); f32[n] sdf in metres
    chunk files continue with   u64 n_known (= dimx*dimy*dimz); u8[n_known]
                                u64 n_color (= dimx*dimy*dimz); u8[n_color][3]
                                u64 n_sem   (= dimx*dimy*dimz); u8[n_sem]
    separate colour files:      u64 dimx, dimy, dimz; u64 n; u8[n][3]       (sparse, same order as the locations)
    separate semantic files:    u64 dimx, dimy, dimz; u8[dimz*dimy*dimx]
"""
import numpy as np

_HEADER = np.dtype([("dims", "<u8", 3), ("voxelsize", "<f4"), ("world2grid", "<f4", 16)])


class _Cursor:
    def __init__(self, buf):
        self.buf, self.off = buf, 0

    def take(self, dtype, count=1):
        dtype = np.dtype(dtype)
        nbytes = dtype.itemsize * int(count)
        if self.off + nbytes > len(self.buf):
            raise EOFError("chunk file truncated")
        out = np.frombuffer(self.buf, dtype=dtype, count=int(count), offset=self.off)
        self.off += nbytes
        return out

    def u64(self):
        return int(self.take("<u8")[0])


def sparse_to_dense_np(locs, values, dimx, dimy, dimz, default_val):
    """data_util.sparse_to_dense_np (data_util.py:45-53); locs in z,y,x order."""
    nf = 1 if values.ndim == 1 else values.shape[1]
    dense = np.full([dimz, dimy, dimx, nf], default_val, dtype=values.dtype)
    dense[locs[:, 0], locs[:, 1], locs[:, 2], :] = values.reshape(values.shape[0], nf)
    return dense if nf > 1 else dense.reshape([dimz, dimy, dimx])


def load_sdf(file, load_sparse, load_known, load_color, is_sparse_file=True, color_file=None, load_semantic=False,
             sem_file=None):
    """Same contract as the reference's ``load_sdf`` (data_util.py:64-159), including its return-shape quirks:
    ``load_semantic`` always returns the dense 5-tuple; ``load_sparse`` returns ``([locs, sdf], [dimz,dimy,dimx],
    world2grid, known, color, semantic)``; a file that cannot be read returns five Nones."""
    assert (not load_sparse and not load_known) or (load_sparse != load_known)
    assert (not load_sparse and not load_semantic) or (load_sparse != load_semantic)
    try:
        with open(file, "rb") as f:
            cur = _Cursor(f.read())
        head = cur.take(_HEADER)[0]
    except (OSError, EOFError):
        print("failed to read file:", file)
        return None, None, None, None, None
    dimx, dimy, dimz = (int(v) for v in head["dims"])
    voxelsize = np.float32(head["voxelsize"])
    world2grid = np.array(head["world2grid"], dtype=np.float32).reshape(4, 4)
    if not is_sparse_file:
        raise NotImplementedError("dense .sdf files are not implemented by the reference either (data_util.py:89)")
    num = cur.u64()
    locs = cur.take("<u4", num * 3).astype(np.int32).reshape(num, 3)[:, ::-1].copy()   # x,y,z on disk -> z,y,x
    sdf = cur.take("<f4", num).astype(np.float32)      # a copy: the buffer is read-only
    sdf /= voxelsize
    cells = dimx * dimy * dimz
    known, num_known = None, 0
    if load_color and color_file is None:               # chunk file: the known grid precedes the colours
        num_known = cur.u64()
    if load_known or num_known > 0:
        assert num_known == cells, "known grid has %d entries for %d cells" % (num_known, cells)
        raw = cur.take("u1", num_known)
        if load_known:
            known = raw.reshape(dimz, dimy, dimx).copy()
            near = (sdf >= -1) & (sdf <= 1)
            known[locs[near, 0], locs[near, 1], locs[near, 2]] = 1
            far = sdf > 1
            known[locs[far, 0], locs[far, 1], locs[far, 2]] = 0
    color = None
    if load_color:
        if color_file is not None:
            with open(color_file, "rb") as f:
                ccur = _Cursor(f.read())
            cdims = ccur.take("<u8", 3)
            assert tuple(int(v) for v in cdims) == (dimx, dimy, dimz)
            n = ccur.u64()
            sparse = ccur.take("u1", n * 3).reshape(n, 3)
            color = sparse_to_dense_np(locs, sparse, dimx, dimy, dimz, 0)
        else:
            num_color = cur.u64()
            assert num_color == cells
            color = cur.take("u1", num_color * 3).reshape(dimz, dimy, dimx, 3).copy()
    semantic = None
    if load_semantic:
        if sem_file is not None:
            with open(sem_file, "rb") as f:
                scur = _Cursor(f.read())
            sdims = scur.take("<u8", 3)
            assert tuple(int(v) for v in sdims) == (dimx, dimy, dimz)
            semantic = scur.take("u1", cells).reshape(dimz, dimy, dimx).copy()
        else:
            num_sem = cur.u64()
            assert num_sem == cells
            semantic = cur.take("u1", num_sem).reshape(dimz, dimy, dimx).copy()
        dense = sparse_to_dense_np(locs, sdf[:, np.newaxis], dimx, dimy, dimz, -float("inf"))
        return dense, world2grid, known, color, semantic
    if load_sparse:
        return [locs, sdf], [dimz, dimy, dimx], world2grid, known, color, semantic
    dense = sparse_to_dense_np(locs, sdf[:, np.newaxis], dimx, dimy, dimz, -float("inf"))
    return dense, world2grid, known, color, semantic


def write_chunk_file(path, dims_xyz, voxelsize, world2grid, locs_xyz, sdf_metres, known=None, color=None, semantic=None):
    """Write a chunk file in the layout above (what ``datagen`` produces); used by tests and to feed synthetic data through
    the reference's data path.  ``known`` (dimz,dimy,dimx) u8, ``color`` (dimz,dimy,dimx,3) u8, ``semantic`` (dimz,dimy,dimx)
    u8 are appended in that order when given (a later block requires the earlier ones)."""
    with open(path, "wb") as f:
        np.asarray(dims_xyz, dtype="<u8").tofile(f)
        np.asarray([voxelsize], dtype="<f4").tofile(f)
        np.asarray(world2grid, dtype="<f4").reshape(16).tofile(f)
        np.asarray([len(sdf_metres)], dtype="<u8").tofile(f)
        np.ascontiguousarray(locs_xyz, dtype="<u4").tofile(f)
        np.ascontiguousarray(sdf_metres, dtype="<f4").tofile(f)
        for block in (known, color, semantic):
            if block is None:
                break
            block = np.ascontiguousarray(block, dtype=np.uint8)
            np.asarray([block.size // (3 if block.ndim == 4 else 1)], dtype="<u8").tofile(f)
            block.tofile(f)

"""numpy front-end of the CPU oracle (oracle/raycast_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "raycast_oracle.c")
LIB = os.path.join(HERE, "_build", "libraycast_oracle.so")


def build(force=False):
    """gcc -O2 -ffp-contract=off (the source spells every fma): contraction is the oracle's to decide."""
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-mavx2", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
           "-std=c11", SRC, "-o", LIB, "-lm"]
    subprocess.check_call(cmd)
    return LIB


class Params(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int), ("height", ctypes.c_int),
        ("depth_min", ctypes.c_float), ("depth_max", ctypes.c_float),
        ("thresh_sample_dist", ctypes.c_float), ("ray_increment", ctypes.c_float),
        ("dimx", ctypes.c_int), ("dimy", ctypes.c_int), ("dimz", ctypes.c_int),
        ("num_chunks", ctypes.c_int), ("views_per_chunk", ctypes.c_int), ("max_pixels_per_voxel", ctypes.c_int),
        ("num_locs", ctypes.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def max_threads():
    return int(lib().oracle_max_threads())


def make_params(dims_zyx, width, height, depth_min, depth_max, thresh, inc, num_chunks, views=1, max_pixels=64,
                num_locs=0):
    return Params(int(width), int(height), float(depth_min), float(depth_max), float(thresh), float(inc),
                  int(dims_zyx[2]), int(dims_zyx[1]), int(dims_zyx[0]), int(num_chunks), int(views), int(max_pixels),
                  int(num_locs))


def build_index(locs, num_chunks, dims_zyx):
    locs = _c(locs, np.int64)
    sm = np.empty((num_chunks,) + tuple(dims_zyx), dtype=np.int32)
    lib().oracle_build_index(_p(locs), ctypes.c_int64(locs.shape[0]), _p(sm), int(num_chunks), int(dims_zyx[0]),
                             int(dims_zyx[1]), int(dims_zyx[2]))
    return sm


def raycast_forward(params, sparse_mapping, vals_sdf, vals_color, vals_normal, vals_semantic, view_matrix, intrinsics,
                    threads=1):
    """Returns dict(color, depth, normal, semantic, mapping3dto2d, mapping3dto2d_num, hit_index)."""
    images = params.num_chunks * params.views_per_chunk
    h, w = params.height, params.width
    rows = params.views_per_chunk * params.num_locs
    out = dict(color=np.empty((images, h, w, 3), np.float32), depth=np.empty((images, h, w), np.float32),
               normal=np.empty((images, h, w, 3), np.float32), semantic=np.empty((images, h, w, 14), np.float32),
               mapping3dto2d=np.empty((max(rows, 1), params.max_pixels_per_voxel), np.int32),
               mapping3dto2d_num=np.empty((max(rows, 1),), np.int32), hit_index=np.empty((images, h, w), np.int32))
    sm = _c(sparse_mapping, np.int32)
    a = [_c(x, np.float32) for x in (vals_sdf, vals_color, vals_normal, vals_semantic, view_matrix, intrinsics)]
    lib().oracle_raycast_forward(ctypes.byref(params), _p(sm), *[_p(x) for x in a], _p(out["color"]), _p(out["depth"]),
                                 _p(out["normal"]), _p(out["semantic"]), _p(out["mapping3dto2d"]),
                                 _p(out["mapping3dto2d_num"]), _p(out["hit_index"]), int(threads))
    return out


def raycast_backward(params, grad_color, grad_depth, grad_normal, grad_semantic, sparse_mapping, mapping3dto2d,
                     mapping3dto2d_num):
    """Returns (d_color (N,3), d_depth (N,1), d_normal (N,3), d_semantic (N,14))."""
    n = max(int(params.num_locs), 1)
    d = [np.empty((n, c), np.float32) for c in (3, 1, 3, 14)]
    g = [_c(x, np.float32) for x in (grad_color, grad_depth, grad_normal, grad_semantic)]
    sm, m, mn = _c(sparse_mapping, np.int32), _c(mapping3dto2d, np.int32), _c(mapping3dto2d_num, np.int32)
    lib().oracle_raycast_backward(ctypes.byref(params), *[_p(x) for x in g], _p(sm), _p(m), _p(mn),
                                  *[_p(x) for x in d])
    k = int(params.num_locs)
    return tuple(x[:k] for x in d)


def raycast_occ(params, occ3d, view_matrix, intrinsics):
    occ3d = _c(occ3d, np.uint8)
    vm, ik = _c(view_matrix, np.float32), _c(intrinsics, np.float32)
    out = np.empty((params.num_chunks, 1, params.height, params.width), np.uint8)
    lib().oracle_raycast_occ(ctypes.byref(params), _p(occ3d), _p(vm), _p(ik), _p(out))
    return out

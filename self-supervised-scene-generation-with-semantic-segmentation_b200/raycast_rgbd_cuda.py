"""Drop-in for the reference's native extension module ``raycast_rgbd_cuda``.

Same four entry points, positional tensors, in-place outputs and error behaviour as
``torch/utils/raycast_rgbd/raycast_rgbd_cuda.cpp:57-160`` -- but each is a thin adapter
(tensor -> device pointer, current device + stream) over the C ABI of ``include/spsg_raycast.h``.

    forward(sparse_mapping, locs, vals_sdf, vals_color, vals_normals, vals_semantic, viewMatrixInv,
            imageColor, imageDepth, imageNormal, imageSemantic, mapping3dto2d, mapping3dto2d_num,
            intrinsicParams, opts)
    backward(grad_color, grad_depth, grad_normal, grad_semantic, sparse_mapping, mapping3dto2d,
             mapping3dto2d_num, dims, d_color, d_depth, d_normals, d_semantic)
    construct_dense_sparse_mapping(locs, sparse_mapping)
    raycast_occ(occ3d, occ2d, viewMatrixInv, intrinsicParams, opts)
"""
import ctypes
import weakref

import torch

from . import _native as N

_workspaces = {}


def _check_input(t, name):
    # CHECK_CUDA / CHECK_CONTIGUOUS (raycast_rgbd_cuda.cpp:53-55): AT_ASSERTM -> RuntimeError
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _check_dtype(t, dtype, name):
    if t.dtype != dtype:
        raise RuntimeError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))


def workspace(device, nbytes, owner=None):
    """Scratch tensor of one raycaster (keyed by its ``sparse_mapping`` buffer), grown on demand; the C ABI never
    allocates.  The forward leaves the backward's work list in it, so -- like the reference's ``mapping3dto2d``
    tables (raycast_rgbd.py:29-32) -- it belongs to the module whose buffers the call pair uses.  Entries die with
    their owner's storage."""
    key = (device.type, device.index, None if owner is None else owner.data_ptr())
    entry = _workspaces.get(key)
    if entry is not None and entry[1] is not None and entry[1]() is None and owner is not None:
        entry = None  # the buffer that owned this address is gone; the address now belongs to a new raycaster
    if entry is None or entry[0].numel() < nbytes:
        for k in [k for k, e in _workspaces.items() if e[1] is not None and e[1]() is None]:
            del _workspaces[k]
        ws = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        entry = (ws, None if owner is None else weakref.ref(owner.untyped_storage()))
        _workspaces[key] = entry
    return entry[0]


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (the context manager costs ~10 us of
    host time per call, which matters for a 50 us step)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def _opt_int(v):
    return int(float(v) + 0.5)  # raycast_rgbd_cuda_kernel.cu:445-456


def construct_dense_sparse_mapping(locs, sparse_mapping):
    _check_input(locs, "locs")
    _check_input(sparse_mapping, "sparse_mapping")
    _check_dtype(locs, torch.int64, "locs")
    _check_dtype(sparse_mapping, torch.int32, "sparse_mapping")
    with device_guard(sparse_mapping.device):
        N.check(N.lib.spsg_build_index(N.ptr(locs), locs.shape[0], N.ptr(sparse_mapping), sparse_mapping.shape[0],
                                       sparse_mapping.shape[1], sparse_mapping.shape[2], sparse_mapping.shape[3],
                                       _stream(sparse_mapping.device)))


def _forward_params(sparse_mapping, locs, mapping3dto2d, opts, views_per_chunk=1, flags=0):
    o = opts.tolist() if isinstance(opts, torch.Tensor) else list(opts)
    return N.make_params(width=_opt_int(o[0]), height=_opt_int(o[1]), depth_min=o[2], depth_max=o[3],
                         thresh_sample_dist=o[4], ray_increment=o[5], dimx=_opt_int(o[6]), dimy=_opt_int(o[7]),
                         dimz=_opt_int(o[8]), num_chunks=sparse_mapping.shape[0], views_per_chunk=views_per_chunk,
                         max_pixels_per_voxel=mapping3dto2d.shape[1], num_locs=locs.shape[0], flags=flags)


def forward(sparse_mapping, locs, vals_sdf, vals_color, vals_normals, vals_semantic, viewMatrixInv, imageColor,
            imageDepth, imageNormal, imageSemantic, mapping3dto2d, mapping3dto2d_num, intrinsicParams, opts,
            views_per_chunk=1, flags=0, build_index=False, clear_grads=None):
    """Reference signature (``raycast_color_forward``, raycast_rgbd_cuda.cpp:57-91) plus keyword extensions with
    reference-preserving defaults.  ``clear_grads`` = (d_color, d_depth, d_normals, d_semantic) lets the forward's fill
    pass clear the rows the matching ``backward(..., grads_cleared=True)`` will write (needs ``build_index=True``)."""
    for t, name in ((sparse_mapping, "sparse_mapping"), (locs, "locs"), (vals_sdf, "vals_sdf"),
                    (vals_color, "vals_color"), (vals_normals, "vals_normals"), (vals_semantic, "vals_semantic"),
                    (viewMatrixInv, "viewMatrixInv"), (imageColor, "imageColor"), (imageDepth, "imageDepth"),
                    (imageNormal, "imageNormal"), (imageSemantic, "imageSemantic"), (mapping3dto2d, "mapping3dto2d"),
                    (mapping3dto2d_num, "mapping3dto2d_num"), (intrinsicParams, "intrinsicParams")):
        _check_input(t, name)
    _check_dtype(sparse_mapping, torch.int32, "sparse_mapping")
    _check_dtype(locs, torch.int64, "locs")
    _check_dtype(mapping3dto2d, torch.int32, "mapping3dto2d")
    _check_dtype(mapping3dto2d_num, torch.int32, "mapping3dto2d_num")
    for t, name in ((vals_sdf, "vals_sdf"), (vals_color, "vals_color"), (vals_normals, "vals_normals"),
                    (vals_semantic, "vals_semantic"), (viewMatrixInv, "viewMatrixInv"),
                    (intrinsicParams, "intrinsicParams"), (imageColor, "imageColor"), (imageDepth, "imageDepth"),
                    (imageNormal, "imageNormal"), (imageSemantic, "imageSemantic")):
        _check_dtype(t, torch.float32, name)
    p = _forward_params(sparse_mapping, locs, mapping3dto2d, opts, views_per_chunk, flags)
    images = p.num_chunks * p.views_per_chunk
    n = p.num_locs
    if mapping3dto2d_num.numel() < p.views_per_chunk * n or mapping3dto2d.shape[0] < p.views_per_chunk * n:
        raise RuntimeError("mapping3dto2d has %d rows, needs views_per_chunk*N = %d" %
                           (mapping3dto2d.shape[0], p.views_per_chunk * n))
    if viewMatrixInv.numel() < images * 16 or intrinsicParams.numel() < images * 4:
        raise RuntimeError("viewMatrixInv / intrinsicParams hold fewer than %d images" % images)
    px = images * p.width * p.height
    if imageDepth.numel() < px or imageColor.numel() < 3 * px or imageNormal.numel() < 3 * px or \
            imageSemantic.numel() < 14 * px:
        raise RuntimeError("image buffers are smaller than %d x %d x %d" % (images, p.height, p.width))
    if vals_sdf.numel() < n or vals_color.numel() < 3 * n or vals_normals.numel() < 3 * n or \
            vals_semantic.numel() < 14 * n:
        raise RuntimeError("voxel value tensors hold fewer than N = %d rows" % n)
    dev = vals_sdf.device
    with device_guard(dev):
        nbytes = N.workspace_bytes(p)
        ws = workspace(dev, nbytes, sparse_mapping)
        args = (ctypes.byref(p), N.ptr(sparse_mapping), N.ptr(locs), N.ptr(vals_sdf), N.ptr(vals_color),
                N.ptr(vals_normals), N.ptr(vals_semantic), N.ptr(viewMatrixInv), N.ptr(intrinsicParams),
                N.ptr(imageColor), N.ptr(imageDepth), N.ptr(imageNormal), N.ptr(imageSemantic),
                N.ptr(mapping3dto2d), N.ptr(mapping3dto2d_num))
        if build_index:
            gb = None
            if clear_grads is not None:
                for t, rows in zip(clear_grads, (3, 1, 3, 14)):
                    _check_input(t, "clear_grads")
                    if t.numel() < n * rows:
                        raise RuntimeError("d_* buffers hold fewer than N = %d rows" % n)
                gb = ctypes.byref(N.grad_buffers(*clear_grads))
            N.check(N.lib.spsg_raycast_forward_indexed(*args, gb, N.ptr(ws), ws.numel(), _stream(dev)))
        else:
            if clear_grads is not None:
                raise RuntimeError("clear_grads needs build_index=True")
            N.check(N.lib.spsg_raycast_forward(*args, N.ptr(ws), ws.numel(), _stream(dev)))
    return ws


def backward(grad_color, grad_depth, grad_normal, grad_semantic, sparse_mapping, mapping3dto2d, mapping3dto2d_num,
             dims, d_color, d_depth, d_normals, d_semantic, views_per_chunk=1, grads_cleared=False):
    """Reference signature (``raycast_color_backward``, raycast_rgbd_cuda.cpp:102-140).
    ``dims`` = int32 CPU tensor (or sequence) [batch, Dx, Dy, Dz, N] (raycast_rgbd.py:30-31)."""
    for t, name in ((grad_color, "grad_color"), (grad_depth, "grad_depth"), (grad_normal, "grad_normal"),
                    (grad_semantic, "grad_semantic"), (sparse_mapping, "sparse_mapping"),
                    (mapping3dto2d, "mapping3dto2d"), (mapping3dto2d_num, "mapping3dto2d_num"),
                    (d_color, "d_color"), (d_depth, "d_depth"), (d_normals, "d_normals"), (d_semantic, "d_semantic")):
        _check_input(t, name)
    d = dims.tolist() if isinstance(dims, torch.Tensor) else list(dims)
    n = int(d[4])
    if d_color.shape[0] < n or d_depth.shape[0] < n or d_normals.shape[0] < n or d_semantic.shape[0] < n:
        raise RuntimeError("d_* buffers hold fewer than N = %d rows" % n)
    p = N.make_params(width=grad_color.shape[2], height=grad_color.shape[1], depth_min=0, depth_max=0,
                      thresh_sample_dist=0, ray_increment=0, dimx=int(d[1]), dimy=int(d[2]), dimz=int(d[3]),
                      num_chunks=sparse_mapping.shape[0], views_per_chunk=views_per_chunk,
                      max_pixels_per_voxel=mapping3dto2d.shape[1], num_locs=n,
                      flags=N.SPSG_FLAG_GRADS_CLEARED if grads_cleared else 0)
    dev = grad_color.device
    with device_guard(dev):
        ws = workspace(dev, N.workspace_bytes(p), sparse_mapping)
        N.check(N.lib.spsg_raycast_backward(ctypes.byref(p), N.ptr(grad_color), N.ptr(grad_depth), N.ptr(grad_normal),
                                            N.ptr(grad_semantic), N.ptr(sparse_mapping), N.ptr(mapping3dto2d),
                                            N.ptr(mapping3dto2d_num), N.ptr(d_color), N.ptr(d_depth),
                                            N.ptr(d_normals), N.ptr(d_semantic), N.ptr(ws), ws.numel(), _stream(dev)))


def raycast_occ(occ3d, occ2d, viewMatrixInv, intrinsicParams, opts, flags=0):
    """Reference signature (``raycast_occ_forward``, raycast_rgbd_cuda.cpp:142-153);
    opts = [W, H, depth_min, depth_max, ray_increment, ...] (raycast_rgbd_cuda_kernel.cu:596-602)."""
    for t, name in ((occ3d, "occ3d"), (occ2d, "occ2d"), (viewMatrixInv, "viewMatrixInv"),
                    (intrinsicParams, "intrinsicParams")):
        _check_input(t, name)
    _check_dtype(occ3d, torch.uint8, "occ3d")
    _check_dtype(occ2d, torch.uint8, "occ2d")
    o = opts.tolist() if isinstance(opts, torch.Tensor) else list(opts)
    p = N.make_params(width=_opt_int(o[0]), height=_opt_int(o[1]), depth_min=o[2], depth_max=o[3],
                      thresh_sample_dist=0, ray_increment=o[4], dimx=occ3d.shape[4], dimy=occ3d.shape[3],
                      dimz=occ3d.shape[2], num_chunks=occ3d.shape[0], flags=flags)
    dev = occ3d.device
    with device_guard(dev):
        N.check(N.lib.spsg_raycast_occ(ctypes.byref(p), N.ptr(occ3d), N.ptr(occ2d), N.ptr(viewMatrixInv),
                                       N.ptr(intrinsicParams), _stream(dev)))

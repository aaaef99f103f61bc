"""Dense generator heads -> the raycaster's sparse inputs (reference ``torch/train.py:494-509``; SURVEY.md section 8(f)
rank 1).

The reference builds the voxel list and its payloads with PyTorch indexing::

    locs = torch.nonzero((torch.abs(output_sdf.detach()[:, 0]) < args.truncation) & ~empty[:, 0])    # train.py:495
    locs = torch.cat([locs[:, 1:], locs[:, :1]], 1)                                                    # :498
    output_sdf = [locs, output_sdf[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]]                # :499
    output_color = [locs, output_color[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]]            # :505
    output_semantic = output_semantic[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]              # :508

``sparse_locs`` returns the same ``locs`` (same rows, same order, int64, columns z, y, x, b) from one ordered stream
compaction, and ``gather_dense`` the same values for any number of heads from fused gather launches, with the matching
backward (zero-filled dense gradient + scatter).  No CPU path.
"""

import torch
from torch.autograd import Function

from . import _native as N
from . import raycast_rgbd_cuda as rc

_MAX_PAYLOADS = 4
_scratch = {}


def _scratch_for(device, nbytes):
    key = (device.type, device.index)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def sparse_locs(sdf, truncation, empty=None):
    """``locs`` (N,4) int64 rows (z, y, x, b) of the voxels with ``|sdf| < truncation`` (and ``~empty``), in
    ``torch.nonzero`` order.  sdf / empty: (B,Dz,Dy,Dx) or (B,1,Dz,Dy,Dx); empty is a bool mask or None."""
    if sdf.dim() == 5:
        if sdf.shape[1] != 1:
            raise RuntimeError("sdf must have one channel")
        sdf = sdf[:, 0]
    if sdf.dim() != 4:
        raise RuntimeError("sdf must be (B,Dz,Dy,Dx) or (B,1,Dz,Dy,Dx)")
    sdf = sdf.detach()
    rc._check_input(sdf, "sdf")
    rc._check_dtype(sdf, torch.float32, "sdf")
    if empty is not None:
        if empty.dim() == 5:
            empty = empty[:, 0]
        if empty.shape != sdf.shape:
            raise RuntimeError("empty must have the shape of sdf")
        if empty.dtype not in (torch.bool, torch.uint8):
            raise RuntimeError("empty must be a bool (or uint8) mask")
        rc._check_input(empty, "empty")
    if sdf.data_ptr() % 16:
        sdf = sdf.clone()  # the kernels read 16-byte vectors
    if empty is not None and empty.data_ptr() % 8:
        empty = empty.clone()
    dev = sdf.device
    B, dz, dy, dx = sdf.shape
    cells = sdf.numel()
    if cells == 0:
        return torch.zeros(0, 4, dtype=torch.int64, device=dev)
    with rc.device_guard(dev):
        scratch = _scratch_for(dev, N.lib.spsg_sparsify_scratch_bytes(cells))
        total = torch.empty(1, dtype=torch.int64, device=dev)
        stream = rc._stream(dev)
        N.check(N.lib.spsg_sparsify_count(N.ptr(sdf), N.ptr(empty), cells, float(truncation), N.ptr(scratch),
                                          scratch.numel(), N.ptr(total), stream))
        n = int(total.item())  # the host needs N to size the outputs (torch.nonzero synchronises for the same reason)
        locs = torch.empty(n, 4, dtype=torch.int64, device=dev)
        N.check(N.lib.spsg_sparsify_locs(N.ptr(sdf), N.ptr(empty), B, dz, dy, dx, float(truncation), N.ptr(scratch),
                                         N.ptr(locs), n, stream))
    return locs


def _payload_array(dense, sparse):
    arr = (N.DensePayload * len(dense))()
    for k, (d, s) in enumerate(zip(dense, sparse)):
        arr[k] = N.DensePayload(d.data_ptr(), s.data_ptr(), d.shape[1], 0)
    return arr


def _run(fn, dense, sparse, locs, shape):
    dev = locs.device
    B, dz, dy, dx = shape
    with rc.device_guard(dev):
        for k0 in range(0, len(dense), _MAX_PAYLOADS):
            d, s = dense[k0:k0 + _MAX_PAYLOADS], sparse[k0:k0 + _MAX_PAYLOADS]
            N.check(fn(_payload_array(d, s), len(d), N.ptr(locs), locs.shape[0], B, dz, dy, dx, rc._stream(dev)))


class _GatherDense(Function):
    @staticmethod
    def forward(ctx, locs, *dense):
        shape = (dense[0].shape[0],) + tuple(dense[0].shape[2:])
        for t in dense:
            rc._check_input(t, "dense head")
            rc._check_dtype(t, torch.float32, "dense head")
            if t.dim() != 5 or (t.shape[0],) + tuple(t.shape[2:]) != shape:
                raise RuntimeError("dense heads must be (B,C,Dz,Dy,Dx) tensors over the same grid")
        rc._check_input(locs, "locs")
        rc._check_dtype(locs, torch.int64, "locs")
        n = locs.shape[0]
        out = tuple(torch.empty(n, t.shape[1], device=t.device) for t in dense)
        _run(N.lib.spsg_dense_gather, dense, out, locs, shape)
        ctx.locs, ctx.shape = locs, shape
        ctx.channels = tuple(t.shape[1] for t in dense)
        return out

    @staticmethod
    def backward(ctx, *grads):
        locs, shape = ctx.locs, ctx.shape
        B, dz, dy, dx = shape
        need = ctx.needs_input_grad[1:]
        idx = [k for k, nd in enumerate(need) if nd]
        sparse = [grads[k].to(torch.float32).contiguous() if grads[k] is not None
                  else torch.zeros(locs.shape[0], ctx.channels[k], device=locs.device) for k in idx]
        dense = [torch.empty(B, ctx.channels[k], dz, dy, dx, device=locs.device) for k in idx]
        if idx:
            _run(N.lib.spsg_dense_scatter, dense, sparse, locs, shape)
        out = [None] * len(need)
        for k, d in zip(idx, dense):
            out[k] = d
        return (None,) + tuple(out)


def gather_dense(locs, *dense):
    """``head[locs[:, -1], :, locs[:, 0], locs[:, 1], locs[:, 2]]`` for every head (B,C,Dz,Dy,Dx) -> (N,C), one fused launch
    per four heads; differentiable with respect to the heads."""
    if not dense:
        return ()
    out = _GatherDense.apply(locs, *[t.contiguous() for t in dense])
    return out if len(dense) > 1 else out[0]


def sparsify_predictions(output_sdf, truncation, empty=None, *heads):
    """train.py:494-509 in one call: ``(locs, sdf_values, *head_values)`` for the dense SDF head (B,1,Dz,Dy,Dx) and any
    further heads (colour, semantics, a one-hot target volume, ...)."""
    locs = sparse_locs(output_sdf, truncation, empty)
    vals = gather_dense(locs, output_sdf, *heads)
    if not heads:
        vals = (vals,)
    return (locs,) + tuple(vals)

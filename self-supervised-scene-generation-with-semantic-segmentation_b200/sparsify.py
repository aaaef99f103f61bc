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
def _prepare(sdf, empty):
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
    return sdf, empty


class CountedLocs:
    """Step 1 of the compaction done ahead of time (``count_locs``): ``n`` rows will come out; the per-tile offsets wait in
    a scratch buffer of their own until ``sparse_locs(..., counted=this)`` writes the rows."""

    def __init__(self, sdf, empty, truncation, scratch, n):
        self.sdf, self.empty, self.truncation, self.scratch, self.n = sdf, empty, float(truncation), scratch, n
        self.version = sdf._version


def count_locs(sdf, truncation, empty=None):
    """How many voxels ``sparse_locs`` will return (one host synchronisation, the one ``torch.nonzero`` has), keeping the
    work done: pass the result as ``counted=`` to ``sparse_locs`` / ``sparsify_predictions`` later.  Lets a training step
    decide early (train.py:524-529) and build the rows late -- after other renders on the same raycaster -- so that
    ``raycaster=`` still feeds the prediction render directly."""
    sdf, empty = _prepare(sdf, empty)
    dev = sdf.device
    cells = sdf.numel()
    if cells == 0:
        return CountedLocs(sdf, empty, truncation, None, 0)
    with rc.device_guard(dev):
        scratch = torch.empty(max(int(N.lib.spsg_sparsify_scratch_bytes(cells)), 256), dtype=torch.uint8, device=dev)
        total = torch.empty(1, dtype=torch.int64, device=dev)
        N.check(N.lib.spsg_sparsify_count(N.ptr(sdf), N.ptr(empty), cells, float(truncation), N.ptr(scratch),
                                          scratch.numel(), N.ptr(total), rc._stream(dev)))
        n = int(total.item())  # the host needs N to size the outputs (torch.nonzero synchronises for the same reason)
    return CountedLocs(sdf, empty, truncation, scratch, n)


def sparse_locs(sdf, truncation, empty=None, raycaster=None, counted=None):
    """``locs`` (N,4) int64 rows (z, y, x, b) of the voxels with ``|sdf| < truncation`` (and ``~empty``), in
    ``torch.nonzero`` order.  sdf / empty: (B,Dz,Dy,Dx) or (B,1,Dz,Dy,Dx); empty is a bool mask or None.

    ``raycaster`` (a ``RaycastRGBD`` over the same grid): the pass that writes ``locs`` also writes the raycaster's voxel
    index (``sparse_mapping``) and dense SDF brick for these rows, so the forward that renders the returned tensor with
    the SDF values gathered from ``sdf`` skips its own fill and index passes (nothing goes through int64 ``locs`` twice).
    The caller's promise: the ``vals_sdf`` passed with these ``locs`` are ``gather_dense(locs, sdf)``.
    ``counted``: the result of ``count_locs`` on the same (unmodified) tensors."""
    if counted is None:
        counted = count_locs(sdf, truncation, empty)
    else:
        s2, _ = _prepare(sdf, empty)
        if (s2.shape != counted.sdf.shape or s2.data_ptr() != counted.sdf.data_ptr() or s2._version != counted.version or
                float(truncation) != counted.truncation):
            raise RuntimeError("counted= belongs to another sdf tensor / truncation (or the tensor was modified since)")
    sdf, empty, n, scratch = counted.sdf, counted.empty, counted.n, counted.scratch
    dev = sdf.device
    B, dz, dy, dx = sdf.shape
    if sdf.numel() == 0:
        return torch.zeros(0, 4, dtype=torch.int64, device=dev)
    with rc.device_guard(dev):
        stream = rc._stream(dev)
        locs = torch.empty(n, 4, dtype=torch.int64, device=dev)
        if raycaster is not None and _feeds(raycaster, sdf, n):
            m = raycaster
            # the workspace of the forward to come, sized for the most views the module renders (the brick sits at its start)
            p = N.make_params(m.width, m.height, m.depth_min, m.depth_max, m.thresh_sample_dist, m.ray_increment,
                              m.dims3d[2], m.dims3d[1], m.dims3d[0], m.batch_size, m.max_num_frames,
                              m.mapping3dto2d.shape[1], n, 0)
            ws = m.workspace.get(dev, N.workspace_bytes(p))
            N.check(N.lib.spsg_sparsify_locs_indexed(N.ptr(sdf), N.ptr(empty), B, dz, dy, dx, float(truncation),
                                                     N.ptr(scratch), N.ptr(locs), n, N.ptr(m.sparse_mapping), N.ptr(ws),
                                                     stream))
            m.workspace.mark_prebuilt(locs)
        else:
            N.check(N.lib.spsg_sparsify_locs(N.ptr(sdf), N.ptr(empty), B, dz, dy, dx, float(truncation), N.ptr(scratch),
                                             N.ptr(locs), n, stream))
    return locs


def _feeds(m, sdf, n):
    """can the index + brick of raycaster ``m`` be written for this head: same grid, same device, rows within its buffers"""
    ws = getattr(m, "workspace", None)
    return (ws is not None and n > 0 and tuple(sdf.shape) == (m.batch_size,) + tuple(m.dims3d) and
            m.sparse_mapping.device == sdf.device and tuple(m.sparse_mapping.shape) == tuple(sdf.shape) and
            n * m.max_num_frames <= m.mapping3dto2d.shape[0])


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


def sparsify_predictions(output_sdf, truncation, empty=None, *heads, raycaster=None, counted=None):
    """train.py:494-509 in one call: ``(locs, sdf_values, *head_values)`` for the dense SDF head (B,1,Dz,Dy,Dx) and any
    further heads (colour, semantics, a one-hot target volume, ...).  ``raycaster`` / ``counted``: see ``sparse_locs`` --
    the rows feed that raycaster's next forward directly."""
    locs = sparse_locs(output_sdf, truncation, empty, raycaster=raycaster, counted=counted)
    vals = gather_dense(locs, output_sdf, *heads)
    if not heads:
        vals = (vals,)
    return (locs,) + tuple(vals)

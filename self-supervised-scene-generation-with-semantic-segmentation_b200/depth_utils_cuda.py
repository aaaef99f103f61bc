"""Drop-in for the reference's native extension module ``depth_utils_cuda`` (torch/utils/depth_utils/
depth_utils_cuda.cpp:80-85): the same four entry points with the same positional tensors and in-place outputs, each a thin
adapter over the C ABI of ``include/spsg_raycast.h`` (kernels in csrc/spsg_depth.cu).  No CPU path."""
import torch

from . import _native as N
from .raycast_rgbd_cuda import _check_dtype, _check_input, _stream


def _image(t, name):
    _check_input(t, name)   # CHECK_CUDA / CHECK_CONTIGUOUS, depth_utils_cuda.cpp:27-29
    _check_dtype(t, torch.float32, name)
    if t.dim() != 4:
        raise RuntimeError("%s must be 4-dimensional" % name)


def bilateral_filter_floatmap(output, input, sigmaD, sigmaR):
    """(B,1,H,W) -> (B,1,H,W); depth_utils_cuda.cpp:31-39."""
    _image(output, "output"); _image(input, "input")
    b, _, h, w = input.shape
    with torch.cuda.device(input.device):
        N.check(N.lib.spsg_depth_bilateral_filter(N.ptr(input), N.ptr(output), b, h, w, float(sigmaD), float(sigmaR),
                                                  _stream(input.device)))


def median_fill_depthmap(output, input):
    """(B,1,H,W) -> (B,1,H,W); depth_utils_cuda.cpp:41-47."""
    _image(output, "output"); _image(input, "input")
    b, _, h, w = input.shape
    with torch.cuda.device(input.device):
        N.check(N.lib.spsg_depth_median_fill(N.ptr(input), N.ptr(output), b, h, w, _stream(input.device)))


def convert_depth_to_cameraspace(output, input, intrinsics, depthMin, depthMax):
    """depth (B,1,H,W) -> camera space (B,H,W,3); depth_utils_cuda.cpp:50-60 (depthMin/depthMax are unused there too)."""
    _image(output, "output"); _image(input, "input")
    _check_input(intrinsics, "intrinsics")
    b, _, h, w = input.shape
    with torch.cuda.device(input.device):
        N.check(N.lib.spsg_depth_to_cameraspace(N.ptr(input), N.ptr(intrinsics), N.ptr(output), b, h, w,
                                                _stream(input.device)))


def compute_normals(output, input):
    """camera space (B,H,W,3) -> normals (B,H,W,3); depth_utils_cuda.cpp:63-69."""
    _image(output, "output"); _image(input, "input")
    b, h, w, _ = input.shape
    with torch.cuda.device(input.device):
        N.check(N.lib.spsg_depth_compute_normals(N.ptr(input), N.ptr(output), b, h, w, _stream(input.device)))

"""Drop-in for the reference's ``torch/utils/depth_utils/depth_utils.py`` (the functions of :46-64 and the module
``Depth2Normals`` :66-100) used by train.py:22,142,537,991 to turn the sensor depth frame into target normals and to fill
its holes in place.  SURVEY.md section 8(f) rank 3.

``Depth2Normals.forward`` enqueues the whole pipeline with one native call and synchronises once (to learn whether holes
remain, i.e. whether to return None like the reference); the reference synchronises on ``(depth == 0).any()`` before every
fill round.  Results -- normals, the in-place filled ``depth``, ``filter_helper``, ``camspace`` -- are bit-identical to the
compiled reference extension (tests/test_gpu_depth_utils.py).  No CPU path."""
import torch

from . import _native as N
from . import depth_utils_cuda
from .raycast_rgbd_cuda import _stream


def bilateral_filter_floatmap(filt, img, sigmad, sigmar):
    depth_utils_cuda.bilateral_filter_floatmap(filt, img, sigmad, sigmar)


def median_fill_depthmap(filt, img, num_iters):
    assert num_iters >= 2
    for _ in range(num_iters // 2):   # depth_utils.py:57-59: img <- fill(filt), filt <- fill(img)
        depth_utils_cuda.median_fill_depthmap(img, filt)
        depth_utils_cuda.median_fill_depthmap(filt, img)


def convert_depth_to_cameraspace(camspace, filt, intrinsic, depth_min, depth_max):
    depth_utils_cuda.convert_depth_to_cameraspace(camspace, filt, intrinsic, depth_min, depth_max)


def compute_normals(normals, camspace):
    depth_utils_cuda.compute_normals(normals, camspace)


class Depth2Normals(torch.nn.Module):
    def __init__(self, batch_size, width, height, depth_min, depth_max, max_num_fill_iters=40, device=None):
        super(Depth2Normals, self).__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("Depth2Normals needs a CUDA device: this implementation has no CPU path")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if max_num_fill_iters // 2 > N.SPSG_DEPTH_MAX_FILL_ROUNDS:
            raise RuntimeError("max_num_fill_iters > %d" % (2 * N.SPSG_DEPTH_MAX_FILL_ROUNDS))
        self.width = width
        self.height = height
        self.depth_min = depth_min
        self.depth_max = depth_max
        self.max_num_fill_iters = max_num_fill_iters
        # pre-allocated helpers, as depth_utils.py:75-77
        self.filter_helper = torch.zeros(batch_size, 1, height, width, device=device)
        self.camspace = torch.zeros(batch_size, height, width, 3, device=device)
        self.normals = torch.zeros(batch_size, height, width, 3, device=device)
        self.hole_counts = torch.zeros(N.SPSG_DEPTH_MAX_FILL_ROUNDS + 1, dtype=torch.int32, device=device)

    def get_campos(self):
        return self.camspace

    def forward(self, depth, intrinsic_params):
        """depth (B,1,H,W) float32 metres, 0 = hole -- MODIFIED IN PLACE when it has holes (filled from the bilateral-
        filtered frame), exactly like the reference; returns normals (B,3,H,W) or None if holes remain."""
        if not depth.is_cuda or not depth.is_contiguous() or depth.dtype != torch.float32:
            raise RuntimeError("depth must be a contiguous CUDA float32 tensor")
        if not intrinsic_params.is_cuda or not intrinsic_params.is_contiguous():
            raise RuntimeError("intrinsic_params must be a contiguous CUDA tensor")
        if depth.dim() != 4 or depth.shape[1] != 1:
            raise RuntimeError("depth must be (B,1,H,W)")
        b, _, h, w = depth.shape
        # the native call writes filter_helper / camspace / normals rows for b x h x w pixels: they must fit the buffers
        if b > self.normals.shape[0] or (h, w) != tuple(self.normals.shape[1:3]):
            raise RuntimeError("depth is %dx%dx%d, this Depth2Normals was built for batches of up to %d frames of %dx%d"
                               % (b, h, w, self.normals.shape[0], self.normals.shape[1], self.normals.shape[2]))
        if intrinsic_params.dtype != torch.float32 or intrinsic_params.numel() < 4 * b or \
                intrinsic_params.device != depth.device:
            raise RuntimeError("intrinsic_params must be float32 (B,4) on the device of depth")
        dev = depth.device
        with torch.cuda.device(dev):
            N.check(N.lib.spsg_depth_to_normals(N.ptr(depth), N.ptr(intrinsic_params), N.ptr(self.filter_helper),
                                                N.ptr(self.camspace), N.ptr(self.normals), N.ptr(self.hole_counts), b, h, w,
                                                2.0, 0.1, int(self.max_num_fill_iters), _stream(dev)))
        if self.max_num_fill_iters > 0 and int(self.hole_counts[self.max_num_fill_iters // 2].item()) != 0:
            return None   # depth_utils.py:92-93
        return self.normals[:b].permute(0, 3, 1, 2).contiguous()

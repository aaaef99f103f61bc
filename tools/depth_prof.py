"""Development aid: the Depth2Normals pipeline on 8 frames 320x256 with 3 % holes (for ncu / timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spsg_b200 import synthetic as S
from spsg_b200.depth_utils import Depth2Normals
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
B, W, H = 8, S.WIDTH, S.HEIGHT
g = torch.Generator(device=dev).manual_seed(3)
base = torch.rand(B, 1, H, W, device=dev, generator=g) * 2.0 + 0.5
base[torch.rand(B, 1, H, W, device=dev, generator=g) < 0.03] = 0.0
intr = torch.tensor([list(S.INTRINSICS)] * B, device=dev)
mod = Depth2Normals(B, W, H, S.DEPTH_MIN, S.DEPTH_MAX, device=dev)
for _ in range(3):
    out = mod(base.clone(), intr)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    out = mod(base.clone(), intr)
b.record(); torch.cuda.synchronize()
print("Depth2Normals %d frames: %.0f us per call" % (B, a.elapsed_time(b) / 20 * 1e3), None if out is None else float(out.abs().mean()))

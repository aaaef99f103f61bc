"""Development aid: device time of the forward / gather kernels on C2 and C3 for the library in $SPSG_RAYCAST_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spsg_b200 import synthetic as S, _native as N
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views
dev = torch.device("cuda", 0)
res = []
for B, F in ((1, 1), (8, 5)):
    batch, t = scene_tensors(list(range(B)), dev)
    n = t["locs"].shape[0]
    fw = []
    gw = []
    for seed in (0, 1, 2):
        _, _, view, intr = views(B, F, dev, seed=seed)
        m = RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, max_num_frames=F, max_num_locs_per_sample=(n + B - 1) // B + 1000, device=dev)
        sdf = t["sdf"].clone().requires_grad_(True); sem = t["semantic"].clone().requires_grad_(True)
        out = m(t["locs"], sdf, t["color"], t["normal"], sem, view, intr)
        grads = [torch.randn_like(o) for o in out]
        for it in range(12):
            if it == 2:
                torch.cuda.synchronize(); N.timing_read(0); N.timing_read(1); N.timing_enable(True)
            o = m(t["locs"], sdf, t["color"], t["normal"], sem, view, intr)
            torch.autograd.backward(o, grads)
        torch.cuda.synchronize(); N.timing_enable(False)
        f_ms, f_n = N.timing_read(0); g_ms, g_n = N.timing_read(1)
        fw.append(f_ms / f_n * 1e3); gw.append(g_ms / g_n * 1e3)
    res.append("B=%d F=%d fwd %s us gather %s us" % (B, F, "/".join("%.1f" % x for x in fw), "/".join("%.1f" % x for x in gw)))
print(os.path.basename(os.environ.get("SPSG_RAYCAST_LIB", "default")), " | ".join(res))

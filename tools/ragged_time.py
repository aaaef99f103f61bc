"""Development aid: forward time of a ragged batch (half of the chunks without voxels) next to the balanced one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spsg_b200 import synthetic as S, _native as N
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views
dev = torch.device("cuda", 0)
B, F = 8, 5
_, t = scene_tensors(list(range(B)), dev)
_, _, view, intr = views(B, F, dev, seed=0)
for name, keep in (("balanced", list(range(B))), ("ragged (chunks 0-3 only)", [0, 1, 2, 3]), ("one chunk of 8", [5])):
    m = torch.zeros(t["locs"].shape[0], dtype=torch.bool, device=dev)
    for b in keep:
        m |= t["locs"][:, 3] == b
    d = {k: v[m].contiguous() for k, v in t.items()}
    n = d["locs"].shape[0]
    rc = RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT,
                     max_num_frames=F, max_num_locs_per_sample=n // B + 1000, device=dev)
    with torch.no_grad():
        for it in range(12):
            if it == 2:
                torch.cuda.synchronize(); N.timing_read(0); N.timing_enable(True)
            rc(d["locs"], d["sdf"], d["color"], d["normal"], d["semantic"], view, intr)
    torch.cuda.synchronize(); N.timing_enable(False)
    ms, cnt = N.timing_read(0)
    print("%-28s %8d voxels  forward %.1f us" % (name, n, ms / cnt * 1e3))

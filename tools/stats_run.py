"""Development aid: event statistics of the forward march (needs the -DSPSG_STATS build, see tools/build_stats.sh)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SPSG_RAYCAST_LIB"] = os.path.join(ROOT, "self-supervised-scene-generation-with-semantic-segmentation_b200", "lib", "libspsg_raycast_stats.so")
sys.path.insert(0, ROOT)
import torch
from spsg_b200 import synthetic as S, _native as N
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views
dev = torch.device("cuda", 0)
names = ["exact", "dense", "invalid", "sign", "jumpE", "jumpSame", "", "", "warp_iters", "lane_events", "steps_jumped", "refine_rounds", "refine_lanes", "", "march_lanes"]
for B, F in ((1, 1), (8, 5)):
    batch, t = scene_tensors(list(range(B)), dev)
    n = t["locs"].shape[0]
    _, _, view, intr = views(B, F, dev, seed=0)
    m = RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, max_num_frames=F, max_num_locs_per_sample=(n + B - 1) // B + 1000, device=dev)
    out = (ctypes.c_ulonglong * 48)()
    N.lib.spsg_debug_stats(out, 1)
    for rep in range(3):   # the last (warm) repetition is reported
        with torch.no_grad():
            m(t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view, intr)
        torch.cuda.synchronize()
        N.lib.spsg_debug_stats(out, 1)
    rays = B * F * S.WIDTH * S.HEIGHT
    warps = rays // 32
    print("B=%d F=%d rays=%d" % (B, F, rays))
    for k, nm in enumerate(names):
        if nm:
            print("  %-14s %12d  per ray %.2f  per warp %.2f" % (nm, out[k], out[k] / rays, out[k] / warps))
    for k, nm in ((16, "setup+clip"), (28, "wait maps"), (18, "march"), (20, "refine"), (22, "epilogue"), (40, "  payload+atom"), (42, "  list append"), (44, "  stage+store"), (24, "whole")):
        print("  cycles %-10s avg/warp %8.0f  max %8d" % (nm, out[k] / warps, out[k + 1]))
    if out[13]:
        print("  per warp: wait for TMA after prepare avg %.0f max %d; block-map build (incl. its barriers) avg %.0f max %d" % (out[4] / out[13], out[5], out[6] / out[13], out[7]))
    print("  warp iterations avg %.1f max" % (out[27] / warps), out[26], " histogram (<=8,16,32,...):", [out[32 + i] for i in range(10)])
    cnt = m.mapping3dto2d_num[:n * F]
    hist = torch.bincount(cnt.clamp(max=64))
    print("  pixels/(voxel,view) histogram:", hist[:12].tolist(), ">4:", int((cnt > 4).sum()), "of", int((cnt > 0).sum()))

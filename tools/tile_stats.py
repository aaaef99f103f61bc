"""Development aid (needs the -DSPSG_STATS=2 build): per-tile cycle breakdown of the C2 forward, slowest tiles first."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SPSG_RAYCAST_LIB"] = os.path.join(ROOT, "self-supervised-scene-generation-with-semantic-segmentation_b200", "lib", "libspsg_raycast_stats.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
from spsg_b200 import synthetic as S, _native as N
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views
dev = torch.device("cuda", 0)
batch, t = scene_tensors([0], dev)
n = t["locs"].shape[0]
_, _, view, intr = views(1, 1, dev, seed=0)
m = RaycastRGBD(1, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, max_num_locs_per_sample=n + 1000, device=dev)
for rep in range(3):
    with torch.no_grad():
        out = m(t["locs"], t["sdf"], t["color"], t["normal"], t["semantic"], view, intr)
    torch.cuda.synchronize()
buf = (ctypes.c_int * (131072 * 8))()
N.lib.spsg_debug_tile_stats(buf)
a = np.frombuffer(buf, dtype=np.int32).reshape(131072, 8)[:2560]
hit = (out[1][0] != -float("inf")).cpu().numpy()
order = np.argsort(-a[:, 0])
print("tile  total  setup  march refine epilog iters smid  start | hits")
tiles_x = S.WIDTH // 8
def tile_xy(t):
    blk, sub = t >> 2, t & 3
    by, bx = divmod(blk, (tiles_x + 1) // 2)
    return (bx * 2 + (sub & 1)) * 8, (by * 2 + (sub >> 1)) * 4
for t_ in list(order[:15]) + list(order[1270:1275]) + list(order[-5:]):
    x0, y0 = tile_xy(int(t_))
    print("%4d %6d %6d %6d %6d %6d %5d %4d %6d | %2d  (x0=%d,y0=%d)" % (t_, *a[t_], hit[y0:y0 + 4, x0:x0 + 8].sum(), x0, y0))
t0 = a[:, 7].min()
print("kernel span (cycles, from first tile start to last tile end): %d" % ((a[:, 7] + a[:, 0]).max() - t0))
print("mean total %.0f; start offsets: min %d max %d" % (a[:, 0].mean(), 0, (a[:, 7] - t0).max()))
per_sm = {}
for r in a:
    per_sm.setdefault(int(r[6]), []).append(int(r[0]))
cnt = np.array([len(v) for v in per_sm.values()])
print("tiles per SM: min %d max %d; SMs used %d" % (cnt.min(), cnt.max(), len(per_sm)))
print("histogram of tile totals (k cycles):", np.histogram(a[:, 0] / 1000.0, bins=[0, 15, 20, 25, 30, 35, 40, 50, 60, 70])[0].tolist())
sm_max = sorted(((max(v), np.mean(v), k) for k, v in per_sm.items()), reverse=True)
print("slowest SMs (max tile, mean tile, smid):", [(int(m), int(mean), k) for m, mean, k in sm_max[:12]])
print("fastest SMs:", [(int(m), int(mean), k) for m, mean, k in sm_max[-6:]])
for name, col in (("setup", 1), ("march", 2), ("refine", 3), ("epilogue", 4)):
    print("%-8s mean %6.0f  p50 %6.0f  p90 %6.0f  p99 %6.0f  max %6d" % (name, a[:, col].mean(), np.percentile(a[:, col], 50), np.percentile(a[:, col], 90), np.percentile(a[:, col], 99), a[:, col].max()))

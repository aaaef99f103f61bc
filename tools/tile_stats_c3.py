"""Development aid (needs the -DSPSG_STATS=2 build, tools/build_stats.sh): per-tile timeline of one fused C3 forward --
how long tiles take, when the last ones start, and how much of the launch is tail (warps without a tile)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SPSG_RAYCAST_LIB"] = os.path.join(ROOT, "self-supervised-scene-generation-with-semantic-segmentation_b200", "lib", "libspsg_raycast_stats.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from spsg_b200 import _native as N
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
o = bench.Ours(dev, 0, 8, 5, 1)
for i in range(3):
    o.step_fused(i)
torch.cuda.synchronize()
buf = (ctypes.c_int * (131072 * 8))()
N.lib.spsg_debug_tile_stats(buf)
T = 2560 * 5 * 8
a = np.frombuffer(buf, dtype=np.int32).reshape(131072, 8)[:T].astype(np.int64)
dur, start, smid = a[:, 0], a[:, 7], a[:, 6]
# clock64 is per SM; SM clocks are not synchronised exactly but start within the same few microseconds: use per-SM offsets
# relative to each SM's first tile start
t0 = {s: start[smid == s].min() for s in np.unique(smid)}
rel = start - np.array([t0[s] for s in smid])
rel = np.where(rel < 0, rel + (1 << 31), rel)
end = rel + dur
span = end.max()
print("tiles %d, SMs %d, mean tile %.0f cycles, p50 %.0f p90 %.0f p99 %.0f max %d" % (T, len(t0), dur.mean(), np.percentile(dur, 50), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
print("launch span (first tile start -> last tile end, per-SM clocks): %d cycles = %.1f us at 1.965 GHz" % (span, span / 1965.0))
warps = 148 * 28
print("sum of tile cycles / (warps * span) = %.3f (warp-slot utilisation)" % (dur.sum() / (warps * span)))
per_sm_end = np.array([end[smid == s].max() for s in t0])
print("per-SM last tile end: min %.1f us, mean %.1f, max %.1f" % (per_sm_end.min() / 1965.0, per_sm_end.mean() / 1965.0, per_sm_end.max() / 1965.0))
# how many tiles are still running at time t
for frac in (0.7, 0.8, 0.85, 0.9, 0.95, 0.98):
    t = frac * span
    running = int(((rel <= t) & (end > t)).sum())
    print("at %.0f%% of the span: %d tiles in flight of %d warp slots" % (100 * frac, running, warps))
last_start = rel.max()
print("last tile starts at %.1f us (%.1f%% of the span)" % (last_start / 1965.0, 100.0 * last_start / span))
order = np.argsort(-end)[:12]
print("tiles that end last:  tile chunk view  start_us  dur_us  march refine epil iters")
for t_ in order:
    print("   %6d %2d %2d  %7.1f %7.1f  %6d %6d %6d %4d" % (t_, t_ // 12800, (t_ % 12800) // 2560, rel[t_] / 1965.0, dur[t_] / 1965.0, a[t_, 2], a[t_, 3], a[t_, 4], a[t_, 5]))
print("histogram of tile durations (us):", np.histogram(dur / 1965.0, bins=[0, 5, 10, 15, 20, 30, 40, 60, 80, 120, 200, 1e9])[0].tolist())
pv = [(c, v, dur[c * 12800 + v * 2560:(c * 12800 + (v + 1) * 2560)].mean() / 1965.0) for c in range(8) for v in range(5)]
print("mean tile duration per (chunk, view) in us:", " ".join("%d/%d:%.0f" % x for x in pv))

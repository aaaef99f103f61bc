"""Development aid: where the end-to-end step of bench.py spends its time -- the host->device copy alone, the
render + losses + backward alone (inputs already in the device slots), and both (what bench.py reports as e2e)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:]]
args = bench.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, F = (1, 1) if args.workload == "c2" else (8, 5)
ctx = bench.run_ours(args, dev, 0, B, F, args.sets or 4)
nbytes = bench.bytes_of(ctx["host"][0], bench.H2D_KEYS)
for mode in ("copy", "compute", "full", "copy", "compute", "full"):
    ms, _ = bench.e2e_ours(ctx, dev, 1, 60, 5, mode=mode)
    print("%-8s %.1f us/step  (%.1f GB/s of input)" % (mode, ms / 60 * 1e3, nbytes / (ms / 60 * 1e-3) / 1e9))

"""Development aid: where the end-to-end step of bench.py spends its time -- the host->device copy alone, the
render + losses + backward alone (inputs already in the device slots), and both (what bench.py reports as e2e).
Runs under torchrun too (every rank probes its own GPU at the same time; max over ranks is printed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

args = bench.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    print("cpus: os.cpu_count=%d affinity=%d" % (os.cpu_count(), len(os.sched_getaffinity(0))))
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        print("numa nodes:", nodes, [open("/sys/devices/system/node/%s/cpulist" % n).read().strip() for n in nodes])
    except OSError as e:
        print("numa: n/a", e)
    os.system("nvidia-smi topo -m 2>/dev/null | head -14 | cut -c1-150")
B, F = (1, 1) if args.workload == "c2" else (8, 5)
ctx = bench.run_ours(args, dev, rank, B, F, args.sets or 4)
nbytes = bench.bytes_of(ctx["host"][0], bench.H2D_KEYS)
for mode in ("copy", "compute", "full", "copy", "compute", "full"):
    ms, _ = bench.e2e_ours(ctx, dev, world, 60, 5, mode=mode)
    if rank == 0:
        print("%-8s %.1f us/step  (%.1f GB/s of input per rank)" % (mode, ms / 60 * 1e3, nbytes / (ms / 60 * 1e-3) / 1e9))
if world > 1:
    dist.destroy_process_group()

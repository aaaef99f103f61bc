"""Development aid: the two halves of bench.py's end-to-end step, timed alone -- (a) the host -> device copy of one packed
pinned input set, (b) the public API path (render_with_2d_losses + backward) with the inputs already in HBM.
usage: python tools/e2e_probe.py [--c2]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, F = (1, 1) if "--c2" in sys.argv else (8, 5)
o = bench.Ours(dev, 0, B, F, 2)
S = o.S


def timed(fn, n=40, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


packed = [bench.pack_host(h) for h in o.host]
slot = torch.empty(max(b.numel() for b, _ in packed), dtype=torch.uint8, device=dev)
copy_ms = timed(lambda i: slot[:packed[i % 2][0].numel()].copy_(packed[i % 2][0], non_blocking=True))


def api(i):
    d = o.devsets[i % 2]
    leaves = [d[k].detach().requires_grad_(True) for k in ("sdf", "color", "semantic")]
    total, _, _ = o.render(o.mods[i % 2], d["locs"], leaves[0], leaves[1], d["normal"], leaves[2], d["view"], d["intr"],
                           images_depth=d["t_depth"], images_color=d["t_color"], target2d_label=d["t_label"],
                           weight_semantic_class=o.cw, voxelsize=S.VOXELSIZE)
    total.backward()


api_ms = timed(api)
nbytes = packed[0][0].numel()
print("B=%d F=%d: copy of %.1f MB alone %.3f ms (%.1f GB/s); API path with resident inputs %.3f ms"
      % (B, F, nbytes / 1e6, copy_ms, nbytes / copy_ms / 1e6, api_ms))

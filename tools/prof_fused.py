"""Small fixed workload for ncu: a few fused forward+loss / backward passes of config C3 (or C2 with --c2) through
spsg_b200.losses.render_loss_and_voxel_grads (the path bench.py's `value` times)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, F = (1, 1) if "--c2" in sys.argv else (8, 5)
o = bench.Ours(dev, 0, B, F, 2)
for i in range(4):
    o.step_fused(i)
torch.cuda.synchronize()
print("done", float(o.loss_sink))

"""Development aid: a few steps of bench.py's e2e_train_flow leg (for an ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
o = bench.Ours(dev, 0, 8, 5, 2)
ms, loss, h2d = o.e2e_device_flow(1, 4, 3)
print("flow ms/step", ms / 4, loss)

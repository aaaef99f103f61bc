"""Development aid: the reference generator's forward + backward alone (batch 8, nf 20), default NCDHW layout against
channels_last_3d, and with TF32 convolutions off -- to see what the out-of-scope part of the train step costs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from baseline import ref_loader
from spsg_b200 import synthetic as S
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
torch.backends.cudnn.benchmark = True
model_util = ref_loader.load_module("model")
s = S.make_train_sample(list(range(10, 18)), 1)
inputs = torch.from_numpy(s["input"]).to(dev); mask = torch.from_numpy(s["mask"]).to(dev)
print("cudnn.allow_tf32", torch.backends.cudnn.allow_tf32, "matmul.allow_tf32", torch.backends.cuda.matmul.allow_tf32)

def run(tag, cl, tf32=True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.manual_seed(7)
    model = model_util.Generator(nf_in_geo=1, nf_in_color=4, nf=20, pass_geo_feats=True, truncation=3.0, max_data_size=S.DIMS_ZYX).to(dev).train()
    x, m = inputs, mask
    if cl:
        model = model.to(memory_format=torch.channels_last_3d)
        x = x.contiguous(memory_format=torch.channels_last_3d); m = m.contiguous(memory_format=torch.channels_last_3d)
    def step():
        occ, sdf, col, sem = model(x, m, pred_sdf=[True, True], pred_color=True, pred_semantic=True)
        loss = occ.float().mean() + sdf.float().mean() + col.float().mean() + sem.float().mean()
        loss.backward()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4): l = step()
    b.record(); torch.cuda.synchronize()
    print("%-28s %.1f ms per fwd+bwd   loss %.6f" % (tag, a.elapsed_time(b) / 4, float(l)))

run("NCDHW (default), tf32 conv", False)
run("channels_last_3d, tf32 conv", True)
run("NCDHW, fp32 conv", False, tf32=False)

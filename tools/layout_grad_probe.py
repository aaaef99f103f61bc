"""Development aid: parameter gradients of one train step under {NCDHW, NDHWC} x {TF32, fp32 convolutions} -- is the
difference between the two layouts cuDNN's TF32 rounding (then fp32 agrees) or something in this package's ops?"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from baseline import ref_loader
from spsg_b200 import synthetic as S
from spsg_b200.train_step import ViewGuidedTrainStep, prepare_generator
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
DIMS = (32, 32, 32); W, H = 80, 64
model_util = ref_loader.load_module("model"); loss_util = ref_loader.load_module("loss")
torch.manual_seed(1234)
base = model_util.Generator(nf_in_geo=1, nf_in_color=4, nf=8, pass_geo_feats=True, truncation=3, max_data_size=DIMS).to(dev).train()
sample = S.make_train_sample([0, 1], 1, dims_zyx=DIMS, width=W, height=H, view_kw=dict(center=(16.0, 16.0, 14.0), radius=38.0, height=30.0))
sample = {k: torch.from_numpy(v).to(dev) for k, v in sample.items()}
cw = torch.tensor(S.CLASS_WEIGHTS, dtype=torch.float32, device=dev)

class _Opt:
    def __init__(self, m): self.m = m
    def zero_grad(self, set_to_none=True):
        for p in self.m.parameters(): p.grad = None
    def step(self): pass

res = {}
for layout in ("ncdhw", "ndhwc"):
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        m = copy.deepcopy(base)
        if layout == "ndhwc": prepare_generator(m)
        step = ViewGuidedTrainStep(m, loss_util, 2, DIMS, W, H, cw, max_num_locs_per_sample=int(np.prod(DIMS)), device=dev)
        loss = step({k: v.clone() for k, v in sample.items()}, optimizer=_Opt(m))
        torch.cuda.synchronize()
        g = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])
        res[(layout, tf32)] = (float(loss), step.last["num_locs"], g)
        print(layout, "tf32" if tf32 else "fp32", "loss %.6f" % float(loss), "num_locs", step.last["num_locs"], "|g| %.5f" % float(g.norm()))
keys = list(res)
for i in range(len(keys)):
    for j in range(i + 1, len(keys)):
        a, b = res[keys[i]][2], res[keys[j]][2]
        print(keys[i], "vs", keys[j], "rel grad diff %.5f" % float((a - b).norm() / a.norm()))

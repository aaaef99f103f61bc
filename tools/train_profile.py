"""Development aid: where the time of one train step (bench.py's train leg) goes -- CUDA-event timers around the stages."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from baseline import ref_loader
from spsg_b200 import synthetic as S
from spsg_b200.train_step import ViewGuidedTrainStep
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
torch.backends.cudnn.benchmark = True
model_util = ref_loader.load_module("model"); loss_util = ref_loader.load_module("loss")
torch.manual_seed(7)
model = model_util.Generator(nf_in_geo=1, nf_in_color=4, nf=20, pass_geo_feats=True, truncation=3.0, max_data_size=S.DIMS_ZYX).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
cw = torch.tensor(S.CLASS_WEIGHTS, device=dev)
s = S.make_train_sample(list(range(10, 18)), 1)
sample = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
step = ViewGuidedTrainStep(model, loss_util, 8, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, cw, max_num_locs_per_sample=640000, device=dev)
for i in range(3):
    step(dict(sample, sdf=sample["sdf"].clone()), optimizer=opt)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        step(dict(sample, sdf=sample["sdf"].clone()), optimizer=opt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))

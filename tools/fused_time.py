import os,sys,json
sys.path.insert(0,'/root/repo')
import torch
import bench
from spsg_b200 import _native as N
dev=torch.device('cuda',0); torch.cuda.set_device(dev)
o=bench.Ours(dev,0,8,5,2)
for i in range(6): o.step_fused(i)
torch.cuda.synchronize(); N.timing_read(0); N.timing_read(1); N.timing_enable(True)
for i in range(30): o.step_fused(i)
torch.cuda.synchronize(); N.timing_enable(False)
f=N.timing_read(0); g=N.timing_read(1)
print(os.path.basename(os.environ.get('SPSG_RAYCAST_LIB','default')), 'fwd %.1f us gather %.1f us'%(f[0]/f[1]*1e3, g[0]/g[1]*1e3))

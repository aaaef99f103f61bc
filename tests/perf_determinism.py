import sys; sys.path.insert(0,'/root/repo')
import torch
from spsg_b200 import synthetic as S
from spsg_b200.raycast_rgbd import RaycastRGBD
from tests.common import scene_tensors, views
dev=torch.device('cuda',0)
_, t = scene_tensors([0,1], dev)
n=t['locs'].shape[0]
_,_,view,intr = views(2,1,dev,seed=0)
rc = RaycastRGBD(2, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, max_num_locs_per_sample=n, device=dev)
res=[]
g=None
for rep in range(6):
    leaves=[t[k].clone().requires_grad_(True) for k in ('sdf','color','normal','semantic')]
    out=rc(t['locs'],*leaves,view,intr)
    if g is None:
        gen=torch.Generator(device=dev).manual_seed(1); g=[torch.randn(o.shape,device=dev,generator=gen) for o in out]
    torch.autograd.backward(out,g)
    res.append([x.grad.clone() for x in leaves]+[rc.mapping3dto2d[:n].clone()])
for rep in range(1,6):
    same=[bool(torch.equal(a.view(torch.int32),b.view(torch.int32))) for a,b in zip(res[0][:4],res[rep][:4])]
    print(rep, 'grads bit-identical:', same, 'mapping rows identical:', bool(torch.equal(res[0][4],res[rep][4])), 'max diff', max(float((a-b).abs().max()) for a,b in zip(res[0][:4],res[rep][:4])))

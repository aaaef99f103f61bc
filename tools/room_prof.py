import os, sys, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from spsg_b200 import room as R, synthetic as S, sparsify
from spsg_b200.normals import compute_normals_sparse
from spsg_b200.raycast_rgbd import RaycastRGBD
dev = torch.device('cuda', 0)
dims = (128, 256, 320)
room = R.synthetic_room_sdf(dims, dev)
predict = R.synthetic_predictor(room)
B, F = 8, 5
chunk_dims = (128, 64, 64)
windows = R.chunk_windows(dims)
rc = RaycastRGBD(B, chunk_dims, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, max_num_frames=F, max_num_locs_per_sample=640000, device=dev)
view_np, intr_np = R.window_views(F, chunk_dims)
view = torch.from_numpy(np.tile(view_np, (B, 1, 1))).to(dev); intr = torch.from_numpy(np.tile(intr_np, (B, 1))).to(dev)
grid2cam = torch.inverse(view[::F]).contiguous()
import sys as _s
if '--prepared' in _s.argv:
    predict.prepare_groups([windows[s:s + B] for s in range(0, len(windows), B)], (64, 64))
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); T[name] = T.get(name, 0) + t1 - t0; return t1
for rep in range(3):
    T.clear()
    for s in range(0, len(windows), B):
        group = windows[s:s + B]
        torch.cuda.synchronize(); t = time.perf_counter()
        locs, sdf, color, sem = predict.predict_group(group, (64, 64)); t = tick('predict_group', t)
        edges = torch.searchsorted(locs[:, 3].contiguous(), torch.arange(len(group) + 1, device=dev)).tolist(); t = tick('counts', t)
        with torch.no_grad():
            normals = compute_normals_sparse(locs, sdf, chunk_dims, grid2cam, num_chunks=B); t = tick('normals', t)
            _, depth, _, sem_img = rc(locs, sdf, color, normals, sem, view, intr); t = tick('raycast', t)
            from spsg_b200.losses import labels_from_render
            labels, h = labels_from_render(sem_img, histogram=True); t = tick('labels+hist', t)
print({k: round(v * 1e3, 2) for k, v in T.items()}, 'ms per room; total', round(sum(T.values()) * 1e3, 2))

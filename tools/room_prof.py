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
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); T[name] = T.get(name, 0) + t1 - t0; return t1
for rep in range(3):
    T.clear()
    for s in range(0, len(windows), B):
        group = windows[s:s + B]
        torch.cuda.synchronize(); t = time.perf_counter()
        locs, sdf, color, sem = predict.predict_group(group, (64, 64)); t = tick('predict_group', t)
        counts = torch.bincount(locs[:, 3], minlength=len(group)).tolist(); t = tick('counts', t)
        with torch.no_grad():
            normals = compute_normals_sparse(locs, sdf, chunk_dims, grid2cam, num_chunks=B); t = tick('normals', t)
            _, depth, _, sem_img = rc(locs, sdf, color, normals, sem, view, intr); t = tick('raycast', t)
            labels = R.labels_from_render_logits(sem_img, depth); t = tick('labels', t)
            h = torch.bincount(labels.reshape(-1).long(), minlength=15).to(torch.float64); t = tick('hist', t)
print({k: round(v * 1e3, 2) for k, v in T.items()}, 'ms per room; total', round(sum(T.values()) * 1e3, 2))
# inside predict_group
T.clear()
for s in range(0, len(windows), B):
    group = windows[s:s + B]
    torch.cuda.synchronize(); t = time.perf_counter()
    head = torch.full((len(group), 1, 128, 64, 64), 7.0, device=dev)
    for b, (y0, x0) in enumerate(group):
        win = room[:, y0:y0 + 64, x0:x0 + 64]; head[b, 0, :, :win.shape[1], :win.shape[2]] = win
    t = tick('assemble', t)
    locs, vals = sparsify.sparsify_predictions(head, 3.0); t = tick('sparsify', t)
    origin = torch.tensor(group, dtype=torch.int64, device=dev)
    c, sm = R._payload_from_positions(locs[:, 0], locs[:, 1] + origin[locs[:, 3], 0], locs[:, 2] + origin[locs[:, 3], 1], 0); t = tick('payload', t)
print({k: round(v * 1e3, 2) for k, v in T.items()})

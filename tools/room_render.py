#!/usr/bin/env python
"""BASELINE.json configs[4]: whole-room inference by sliding chunks with multi-view semantic rendering, windows sharded
over the GPUs of one box.  `python tools/room_render.py` (1 GPU) or
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/room_render.py`.
Prints one JSON line (rank 0): rendered windows / s and rays / s, time = max over ranks (CUDA events)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spsg_b200 import parallel as P, room as R, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--room", type=int, nargs=3, default=(128, 256, 320), help="room dims z y x in 2 cm voxels")
ap.add_argument("--views", type=int, default=5)
ap.add_argument("--chunks-per-launch", type=int, default=8)
ap.add_argument("--repeats", type=int, default=10)
args = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
rank, world = P.init_from_env(device=dev)
room = R.synthetic_room_sdf(tuple(args.room), dev)
predict = R.synthetic_predictor(room)
kw = dict(views_per_chunk=args.views, chunks_per_launch=args.chunks_per_launch, rank=rank, world=world)
# the generator is out of scope: its outputs (the dense heads of every launch group of this rank) exist before the clock starts
windows = R.chunk_windows(tuple(args.room))
mine = [windows[i] for i in P.shard_round_robin(len(windows), rank, world)]
if not os.environ.get('SPSG_ROOM_NO_PREPARE'):
    predict.prepare_groups([mine[s:s + args.chunks_per_launch] for s in range(0, len(mine), args.chunks_per_launch)], (64, 64))
out = R.render_room(predict, tuple(args.room), dev, **kw)   # warm-up
P.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.repeats):
    out = R.render_room(predict, tuple(args.room), dev, **kw)
b.record(); torch.cuda.synchronize()
ms = P.max_over_ranks(a.elapsed_time(b) / args.repeats)
rendered = P.sum_over_ranks(out["rendered_windows"]); rays = P.sum_over_ranks(out["rays"])
if rank == 0:
    print(json.dumps({"workload": "room %dx%dx%d, %d windows (64x64 stride 32), %d views 320x256 per window" % (*args.room, out["windows"], args.views),
                      "n_gpus": world, "ms_per_room": ms, "windows_per_s": rendered / (ms * 1e-3), "rays_per_s": rays / (ms * 1e-3),
                      "includes": "from the generator's dense heads on: sparsification (spsg sparsify ops), normals, raycast forward, fused label map + histogram",
                      "label_hist": [int(v) for v in out["label_hist"]]}))
if world > 1:
    torch.distributed.destroy_process_group()

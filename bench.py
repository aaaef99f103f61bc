#!/usr/bin/env python
"""bench.py -- raycast fwd+bwd rays/s on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

A "step" is one forward + backward pass of the hot path over one batch of synthetic input:
  c2 (default, BASELINE configs[1])  one 64x64x128 chunk, one 320x256 view, depth/colour/normal/semantic outputs
  c3 (BASELINE configs[2])           8 chunks x 5 views (40 images, 3 276 800 rays)
`value`   rays/s with every input already resident in HBM, the step replayed from a CUDA graph (no host launch
          latency in the device number); 4+ distinct input sets are rotated so that consecutive steps never find
          their data in the 126 MB L2.
`e2e`     the same metric through the public Python API (RaycastRGBD + fused 2D losses for ours; the reference
          wrapper's call order + its literal PyTorch losses for --impl reference) with HOST buffers: every step
          copies its voxel tensors, cameras and target frames from pinned host memory and reads the loss back.
`roofline`     the dominant kernel (raycast forward) timed with CUDA events on its launch stream (C-ABI timing hook),
               algorithmic bytes 116*Nv + 84*Np per launch over MEASURED_PEAKS.json's HBM copy peak.
`cpu_baseline` the scalar C restatement of the same raycast (oracle/, OpenMP over pixels) on the box's host cores.
--impl reference  runs the UNMODIFIED reference extension (oracle/_ref, built from /root/reference where it lies)
          on the same GPU through its own native entry points in its wrapper's order; the reference has no CPU
          implementation of this path, so its arm is its CUDA extension (falls back to the CPU port if the
          extension is not present on the box).
One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); units are independent chunk x view batches, so
ranks shard them with no data-path collective ("scaling": "weak"); time is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

MAX_LOCS = 640000  # train.py:136 max_num_locs_per_sample (sizes the reference's memsets)
E2E_REPEATS = 3   # timed regions of the end-to-end leg (median reported)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--sets", type=int, default=0, help="distinct resident input sets to rotate (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": reasons}


def make_host_sets(num_sets, B, F, rank):
    """Seeded synthetic chunk batches + cameras + target frames, as pinned host tensors."""
    from spsg_b200 import synthetic as S
    sets = []
    for k in range(num_sets):
        seeds = [1 + rank * 1000 + k * B + b for b in range(B)]
        batch = S.make_batch(seeds)
        view, intr = S.make_views(B, F, seed=rank * 100 + k)
        rng = np.random.default_rng(k + 17 * rank)
        imgs = B * F
        host = {
            "locs": torch.from_numpy(batch["locs"]), "sdf": torch.from_numpy(batch["sdf"]),
            "color": torch.from_numpy(batch["color"]), "normal": torch.from_numpy(batch["normal"]),
            "semantic": torch.from_numpy(batch["semantic"]), "view": torch.from_numpy(view),
            "intr": torch.from_numpy(intr),
            # target frames (2D losses): depth in metres with 5 % holes, colour in [0,1], labels 0..14
            "t_depth": torch.from_numpy(np.where(rng.random((imgs, S.HEIGHT, S.WIDTH)) < 0.05, 0.0,
                                                 rng.uniform(0.8, 1.6, (imgs, S.HEIGHT, S.WIDTH))).astype(np.float32)),
            "t_color": torch.from_numpy(rng.random((imgs, S.HEIGHT, S.WIDTH, 3), dtype=np.float32)),
            "t_label": torch.from_numpy(rng.integers(0, 15, (imgs, S.HEIGHT, S.WIDTH), dtype=np.uint8)),
        }
        sets.append({k2: v.pin_memory() for k2, v in host.items()})
    return sets


def bytes_of(d, keys):
    return int(sum(d[k].numel() * d[k].element_size() for k in keys))


H2D_KEYS = ("locs", "sdf", "color", "normal", "semantic", "view", "intr", "t_depth", "t_color", "t_label")


def run_ours(args, dev, rank, B, F, num_sets):
    from spsg_b200 import _native as N
    from spsg_b200 import synthetic as S
    from spsg_b200.losses import render_with_2d_losses
    from spsg_b200.raycast_rgbd import RaycastRGBD
    host = make_host_sets(num_sets, B, F, rank)
    rays = B * F * S.WIDTH * S.HEIGHT
    devsets, mods, grads = [], [], []
    for h in host:
        d = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
        devsets.append(d)
        mods.append(RaycastRGBD(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST,
                                S.RAY_INCREMENT, max_num_frames=F, max_num_locs_per_sample=MAX_LOCS, device=dev))
        g = torch.Generator(device=dev).manual_seed(5)
        grads.append([torch.randn(s, device=dev, generator=g) for s in
                      ((B * F, S.HEIGHT, S.WIDTH, 3), (B * F, S.HEIGHT, S.WIDTH), (B * F, S.HEIGHT, S.WIDTH, 3),
                       (B * F, S.HEIGHT, S.WIDTH, 14))])
    nv = int(np.mean([d["locs"].shape[0] for d in devsets]))
    cw = torch.tensor(S.CLASS_WEIGHTS, dtype=torch.float32, device=dev)
    from spsg_b200 import raycast_rgbd_cuda as rc

    def step_resident(i):
        d, m, g = devsets[i % num_sets], mods[i % num_sets], grads[i % num_sets]
        n = d["locs"].shape[0]
        opts = [S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, 64, 64, 128]
        rc.forward(m.sparse_mapping, d["locs"], d["sdf"], d["color"], d["normal"], d["semantic"], d["view"],
                   m.image_color, m.image_depth, m.image_normal, m.image_semantic, m.mapping3dto2d,
                   m.mapping3dto2d_num, d["intr"], opts, views_per_chunk=F, build_index=True,
                   clear_grads=(m.d_color, m.d_depth, m.d_normal, m.d_semantic))
        rc.backward(g[0], g[1], g[2], g[3], m.sparse_mapping, m.mapping3dto2d, m.mapping3dto2d_num,
                    [B, 64, 64, 128, n], m.d_color, m.d_depth, m.d_normal, m.d_semantic, views_per_chunk=F,
                    grads_cleared=True)

    # ---- device-resident throughput: the rotation over all input sets captured once into a CUDA graph
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        step_resident(0)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for i in range(num_sets):
                step_resident(i)
    torch.cuda.current_stream(dev).wait_stream(side)
    replays = max(1, (args.steps + num_sets - 1) // num_sets)
    steps = replays * num_sets
    for _ in range(max(1, args.warmup // num_sets)):
        graph.replay()
    torch.cuda.synchronize()
    return dict(host=host, devsets=devsets, mods=mods, rays=rays, nv=nv, graph=graph, replays=replays, steps=steps,
                cw=cw, step_resident=step_resident, render=render_with_2d_losses, N=N, S=S,
                launches_per_step=5)  # fill, index, cell classes, forward, gather


def timed_graph(ctx, dev, world):
    """K steps from the graph, bracketed by barrier + synchronize, CUDA-event time, max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(ctx["replays"]):
        ctx["graph"].replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def e2e_ours(ctx, dev, world, steps, warmup, mode="full"):
    """End to end through the public API with HOST inputs: every step copies its voxel tensors, cameras and target
    frames from pinned host memory (on a copy stream, double-buffered so that step i+1's copy overlaps step i's
    kernels -- what a training loop's prefetcher does), renders + losses + backward, and reads the loss back."""
    S, render, mods, host, cw = ctx["S"], ctx["render"], ctx["mods"], ctx["host"], ctx["cw"]
    num_sets = len(host)
    result = torch.zeros((), pin_memory=True)
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    # Every set is packed into ONE pinned host buffer (256-byte aligned fields, what a loader thread would hand over)
    # and lands in one of two device-side slots with a single async copy; tensors are views into the slot.
    def layout(h):
        off, fields = 0, {}
        for k in H2D_KEYS:
            nbytes = h[k].numel() * h[k].element_size()
            fields[k] = (off, nbytes, h[k].dtype, tuple(h[k].shape))
            off += (nbytes + 255) // 256 * 256
        return fields, off
    packed = []
    for h in host:
        fields, total = layout(h)
        buf = torch.empty(total, dtype=torch.uint8).pin_memory()
        for k, (off, nbytes, dtype, shape) in fields.items():
            buf[off:off + nbytes].view(dtype).view(shape).copy_(h[k])
        packed.append((buf, fields))
    cap = max(b.numel() for b, _ in packed)
    slots = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    pick = (lambda i: i % 2) if mode == "compute" else (lambda i: i % num_sets)  # "compute": the slots keep sets 0/1

    def issue_copy(i):
        slot, (buf, _) = i % 2, packed[pick(i)]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])  # the step that last used this slot is done with it
            slots[slot][:buf.numel()].copy_(buf, non_blocking=True)
            copied[slot].record(copy_stream)

    view_cache = {}

    def views(i):
        # tensor views into a slot depend on (slot, set layout) only: built once, like a loader's collate would
        key = (i % 2, pick(i))
        if key not in view_cache:
            slot, (_, fields) = key[0], packed[key[1]]
            view_cache[key] = {k: slots[slot][off:off + nbytes].view(dtype).view(shape)
                               for k, (off, nbytes, dtype, shape) in fields.items()}
        return view_cache[key]

    def step(i, last):
        slot, m = i % 2, mods[pick(i)]
        if not last and mode != "compute":  # "copy" / "compute": tools/e2e_probe.py times the two halves alone
            issue_copy(i + 1)
        main.wait_event(copied[slot])
        if mode == "copy":
            consumed[slot].record(main)
            return
        d = views(i)
        sdf = d["sdf"].detach().requires_grad_(True)  # fresh leaves every step
        col = d["color"].detach().requires_grad_(True)
        sem = d["semantic"].detach().requires_grad_(True)
        total, terms, _ = render(m, d["locs"], sdf, col, d["normal"], sem, d["view"], d["intr"],
                                 images_depth=d["t_depth"], images_color=d["t_color"], target2d_label=d["t_label"],
                                 weight_semantic_class=cw, voxelsize=S.VOXELSIZE)
        total.backward()
        consumed[slot].record(main)
        result.copy_(total.detach(), non_blocking=True)

    def run(n):
        for e in consumed:
            e.record(main)
        issue_copy(0)
        if mode == "compute":
            issue_copy(1)
        for i in range(n):
            step(i, i == n - 1)

    run(max(3, warmup))
    times = []
    for _ in range(E2E_REPEATS):  # host-side jitter is of the order of the step: median of a few timed regions
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(steps)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        times.append(ms)
    return float(np.median(times)), float(result)


def roofline_ours(ctx, dev, steps):
    """Dominant kernel (raycast forward) alone: CUDA events around each launch on its stream, inputs rotated."""
    N = ctx["N"]
    N.timing_read(0), N.timing_read(1)
    N.timing_enable(True)
    for i in range(steps):
        ctx["step_resident"](i)
    torch.cuda.synchronize()
    N.timing_enable(False)
    f_ms, f_n = N.timing_read(0)
    g_ms, g_n = N.timing_read(1)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, which = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    alg_bytes = 116 * ctx["nv"] + 84 * ctx["rays"]
    us = f_ms / max(f_n, 1) * 1e3
    achieved = alg_bytes / (us * 1e-6) / 1e9 if us > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch, from the committed ncu --set full captures
    # (profiles/r01_c2_ncu_full_summary.md, r01_c3_ncu_full_summary.md); in C2 the 6.9 MB of images are still dirty
    # in the 126 MB L2 when the kernel ends, so only the reads show up
    traffic = {1: 2.52e6, 40: 397.4e6}.get(ctx["rays"] // (ctx["S"].WIDTH * ctx["S"].HEIGHT))
    return {"bound": "hbm", "kernel": "raycast_forward_kernel", "achieved": round(achieved, 1), "peak": peak,
            "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": which,
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_us": round(us, 2),
            "backward_gather_us": round(g_ms / max(g_n, 1) * 1e3, 2),
            "note": "not HBM-bound (traffic <= algorithmic bytes): instruction issue (C3) / dependent latency (C2), see DESIGN.md section 5 and profiles/README.md"}


def cpu_baseline(B, F, seconds=12.0):
    """Scalar C restatement of the raycast fwd+bwd (oracle/raycast_oracle.c), OpenMP over pixels, host cores."""
    from oracle import oracle as O
    from spsg_b200 import synthetic as S
    threads = O.max_threads()
    b = S.make_batch([1])
    view, intr = S.make_views(1, 1, seed=0)
    n = b["locs"].shape[0]
    p = O.make_params(S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST, S.RAY_INCREMENT, 1,
                      1, 64, n)
    rng = np.random.default_rng(0)
    g = None
    t0, steps = time.perf_counter(), 0
    while True:
        sm = O.build_index(b["locs"], 1, S.DIMS_ZYX)
        out = O.raycast_forward(p, sm, b["sdf"], b["color"], b["normal"], b["semantic"], view, intr, threads=threads)
        if g is None:
            g = [rng.standard_normal(out[k].shape).astype(np.float32) for k in ("color", "depth", "normal", "semantic")]
        O.raycast_backward(p, *g, sm, out["mapping3dto2d"], out["mapping3dto2d_num"])
        steps += 1
        dt = time.perf_counter() - t0
        if dt > seconds or steps >= 400:
            break
    rays = S.WIDTH * S.HEIGHT * steps
    return {"value": rays / dt, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": "%d fwd+bwd steps of config c2 (one chunk, one 320x256 view) in %.1f s; forward OpenMP over "
                      "pixels on %d threads, backward scalar" % (steps, dt, threads)}


def run_reference(args, dev, rank, world, B, F, num_sets):
    """The unmodified reference CUDA extension (oracle/_ref) in its wrapper's call order; one view per call as the
    reference renders one view per chunk (F views = F calls)."""
    from oracle import losses_ref as R
    from oracle import ref_driver
    from spsg_b200 import synthetic as S
    host = make_host_sets(num_sets, B, F, rank)
    rays = B * F * S.WIDTH * S.HEIGHT
    cw = torch.tensor(S.CLASS_WEIGHTS, dtype=torch.float32, device=dev)
    ref = ref_driver.RefRaycaster(B, S.DIMS_ZYX, S.WIDTH, S.HEIGHT, S.DEPTH_MIN, S.DEPTH_MAX, S.THRESH_SAMPLE_DIST,
                                  S.RAY_INCREMENT, MAX_LOCS, 64, device=dev)
    devsets = [{k: v.to(dev) for k, v in h.items()} for h in host]
    g = torch.Generator(device=dev).manual_seed(5)
    grads = [torch.randn(s, device=dev, generator=g) for s in
             ((B, S.HEIGHT, S.WIDTH, 3), (B, S.HEIGHT, S.WIDTH), (B, S.HEIGHT, S.WIDTH, 3), (B, S.HEIGHT, S.WIDTH, 14))]
    sel = [torch.arange(B, device=dev) * F + f for f in range(F)]

    def step_resident(i):
        d = devsets[i % num_sets]
        for f in range(F):
            ref.forward(d["locs"], d["sdf"], d["color"], d["normal"], d["semantic"], d["view"][sel[f]].contiguous(),
                        d["intr"][sel[f]].contiguous())
            ref.backward(*grads)

    result = torch.zeros(1, pin_memory=True)

    def step_e2e(i):
        h = host[i % num_sets]
        d = {k: h[k].to(dev, non_blocking=True) for k in H2D_KEYS}
        total = None
        for f in range(F):
            out = ref.forward(d["locs"], d["sdf"], d["color"], d["normal"], d["semantic"],
                              d["view"][sel[f]].contiguous(), d["intr"][sel[f]].contiguous())
            imgs = [o.detach().clone().requires_grad_(True) for o in out]
            label = d["t_label"][sel[f]].unsqueeze(-1)
            loss = R.depth_l1_loss(imgs[1], d["t_depth"][sel[f]].unsqueeze(1), S.VOXELSIZE) + \
                R.compute_2dcolor_loss(imgs[0], d["t_color"][sel[f]], None) + \
                R.semantic_2d_ce_loss(imgs[3], label, cw)
            loss.backward()
            ref.backward(imgs[0].grad, imgs[1].grad, torch.zeros_like(imgs[2]), imgs[3].grad)
            total = loss.detach() if total is None else total + loss.detach()
        result.copy_(total.reshape(1), non_blocking=True)

    def timed(fn, steps, warmup):
        for i in range(max(3, warmup)):
            fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    steps = max(1, min(args.steps, 60 if B * F == 1 else 10))
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms = timed(step_resident, steps, args.warmup)
    sampler.stop_flag = True
    ms_e2e = float(np.median([timed(step_e2e, steps, min(args.warmup, 3)) for _ in range(E2E_REPEATS)]))
    return dict(rays=rays, steps=steps, ms=ms, ms_e2e=ms_e2e, clocks=sampler.summary(),
                h2d=bytes_of(host[0], H2D_KEYS), nv=int(np.mean([d["locs"].shape[0] for d in devsets])))


def main():
    args = parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this implementation has no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, F = (1, 1) if args.workload == "c2" else (8, 5)
    from spsg_b200 import synthetic as S
    per_set_mb = 45.0 * B + 14.0 * B * F
    num_sets = args.sets or max(2, int(np.ceil(190.0 / per_set_mb)))
    config = {"workload": "%s: %d chunk(s) 64x64x128 x %d view(s) 320x256, depth+colour+normal+semantic, fwd+bwd"
                          % (args.workload, B, F),
              "chunks_per_step": B, "views_per_chunk": F, "rays_per_step": B * F * S.WIDTH * S.HEIGHT,
              "max_num_locs_per_sample": MAX_LOCS, "parallelism": "dp%d (independent chunk x view batches)" % world,
              "l2": "rotating %d distinct resident input sets (~%d MB > 126 MB L2) between timed steps"
                    % (num_sets, int(per_set_mb * num_sets))}
    base = {"metric": "raycast fwd+bwd rays/s", "unit": "rays/s", "n_gpus": world, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config}

    if args.impl == "reference":
        from oracle import ref_driver
        if ref_driver.available():
            r = run_reference(args, dev, rank, world, B, F, num_sets)
            value = world * r["rays"] * r["steps"] / (r["ms"] * 1e-3)
            e2e = world * r["rays"] * r["steps"] / (r["ms_e2e"] * 1e-3)
            line = dict(base, impl="reference", value=value, steps=r["steps"], ms_per_step=r["ms"] / r["steps"],
                        clocks=r["clocks"], gpu_launches=0,
                        e2e={"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4},
                        cpu_baseline={"value": value, "unit": "rays/s", "cores": 1, "kind": "reference",
                                      "sample": "reference CUDA extension (oracle/_ref, unmodified sources compiled "
                                                "for sm_100) on the same B200: the reference has no CPU implementation "
                                                "of this path; %d steps, max_num_locs_per_sample=%d as train.py:136"
                                                % (r["steps"], MAX_LOCS)})
        else:
            if rank != 0:
                return
            c = cpu_baseline(B, F)
            line = dict(base, impl="reference", value=c["value"], steps=1, ms_per_step=None, n_gpus=1,
                        clocks={"sm_mhz": None, "sm_max_mhz": None, "reasons": ["cpu run"]}, gpu_launches=0,
                        e2e={"value": c["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                        cpu_baseline=c)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    ctx = run_ours(args, dev, rank, B, F, num_sets)
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms = timed_graph(ctx, dev, world)
    sampler.stop_flag = True
    steps = ctx["steps"]
    value = world * ctx["rays"] * steps / (ms * 1e-3)
    e2e_steps = max(10, min(steps, 100))
    ms_e2e, last_loss = e2e_ours(ctx, dev, world, e2e_steps, min(args.warmup, 5))
    e2e = world * ctx["rays"] * e2e_steps / (ms_e2e * 1e-3)
    roof = roofline_ours(ctx, dev, min(steps, 50)) if rank == 0 else None
    if rank == 0:
        line = dict(base, value=value, steps=steps, ms_per_step=ms / steps, clocks=sampler.summary(),
                    e2e={"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": bytes_of(ctx["host"][0], H2D_KEYS),
                         "d2h_bytes_per_step": 4, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                         "api": "spsg_b200.losses.render_with_2d_losses (fused raycast + depth/colour/semantic losses) "
                                "+ backward; every step's inputs copied from one packed pinned-host buffer on a copy stream (double-buffered); median of %d timed regions" % E2E_REPEATS,
                         "last_loss": last_loss},
                    gpu_launches=ctx["launches_per_step"] * steps, roofline=roof)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(B, F)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

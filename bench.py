#!/usr/bin/env python3
"""Benchmark of Dream2Real's imagination-and-scoring hot path (BASELINE.json metric:
candidate renders + CLIP scores per second at 800x800).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--scene shopping] [--poses 4096] [--res 800] [--clip ViT-B/32] [--chunk 512]

One "step" = one pass of the hot path over one batch of synthetic candidate poses: fused fg render +
depth-test composite -> rot90 + PIL-exact preprocess -> ViT forward on tcgen05 -> score.
N = 1 workload = BASELINE.json configs[1] (shopping scene stand-in, 4096 poses, 800x800, 1 x B200);
N > 1: every rank gets its own 4096 poses (weak scaling) and the ranks exchange scores with one NCCL
all-gather per step.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "candidate renders+CLIP-scores/sec @800x800"
UNIT = "candidates/s"
GOAL = "an apple inside a blue and white bowl"          # reference lang/cache.json (shopping demo)
NORM = ["an apple and a blue and white bowl"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="shopping")
    ap.add_argument("--poses", type=int, default=4096)
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--clip", default="ViT-B/32", choices=["ViT-B/32", "ViT-L/14-336"])
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--log2-hashmap", type=int, default=19)
    ap.add_argument("--cpu-sample", type=int, default=2, help="candidates per CPU-baseline sample / reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def pose_grid(scene, n):
    """First n poses of sample_poses_grid([g, g, 1, 1, 1, 1]) for the scene type (x slowest), g = ceil(sqrt(n))."""
    import types

    import torch

    from dream2real_b200.vision_3d.obj_pose_opt import sample_poses_grid
    g = int(np.ceil(np.sqrt(n)))
    tm = types.SimpleNamespace(scene_model=types.SimpleNamespace(scene_centre=torch.tensor(scene["scene_centre"]), device=torch.device("cpu")))
    res = [g, g, 1, 1, 1, 1] if scene["scene_type"] != 1 else [max(2, int(round(n ** (1 / 3)))) for _ in range(3)] + [1, 1, 1]
    p = sample_poses_grid(tm, res, scene_type=scene["scene_type"])
    reps = int(np.ceil(n / p.shape[0]))
    return p.repeat(reps, 1)[:n].reshape(-1, 4, 4).numpy().astype(np.float64), res


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_world(args, device, scene_dir):
    """Scene, models, CLIP and cached background -- everything that is per query, not per candidate."""
    import torch

    from dream2real_b200 import synth
    from dream2real_b200.clip import ClipVision, make_hf_clip, text_embeds
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    scene = synth.make_scene(args.scene, scene_dir, log2_hashmap_size=args.log2_hashmap, seed=1234)
    tm = synth.SyntheticTaskModel(scene, GOAL, NORM, device)
    rnd = renderer(scene_dir, tm, resolution=args.res, max_candidates_per_launch=args.chunk)
    hf = make_hf_clip(args.clip, seed=1234, vocab_size=49408)
    cv = ClipVision(hf, max_batch=args.chunk, device=device.index)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, 40000, (1 + len(NORM), 12), generator=g)
    ids[:, -1] = 2
    txt = text_embeds(hf, ids).to(device)
    return scene, tm, rnd, hf, cv, txt, accio2ngp


def run_ours(args):
    import torch
    import torch.distributed as dist

    from dream2real_b200 import _native as N
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    scene_dir = tempfile.mkdtemp(prefix=f"d2r_bench_r{rank}_")
    scene, tm, rnd, hf, cv, txt, accio2ngp = build_world(args, device, scene_dir)
    K, res = args.poses, args.res
    poses, grid_res = pose_grid(scene, K)
    # every rank scores its own candidate set (weak scaling): shift the grid a little per rank
    poses[:, 0, 3] += 0.003 * rank
    valid_poses_ngp = accio2ngp.converter(poses)
    render_poses_ngp = accio2ngp.converter(scene["opt_cam_poses"][:1])
    fg = tm.movable_obj.vis_model

    # per-query state resident in HBM: background render + depth, candidate camera matrices
    bg_image, bg_depth = rnd.render_background(render_poses_ngp[0], 0, tm.depths[0], tm.movable_masks[0])
    fg.set_camera_to_training_view(0)
    T1 = accio2ngp.converter(scene["fg_pose"][None])[0]
    cams = T1 @ (np.linalg.inv(valid_poses_ngp) @ T1) @ (np.linalg.inv(T1) @ render_poses_ngp[0])
    cams_ngp = fg.cams_to_ngp(cams[:, :3, :])
    u8 = torch.empty((args.chunk, res, res, 3), dtype=torch.uint8, device=device)
    rects = torch.empty((args.chunk, 4), dtype=torch.int32, device=device)
    bg_u8 = torch.empty((res, res, 3), dtype=torch.uint8, device=device)
    scores = torch.empty(K, dtype=torch.float32, device=device)
    gathered = torch.empty(K * world, dtype=torch.float32, device=device) if world > 1 else None

    vit_events = []      # (start, end) CUDA events around every ViT forward of the timed region

    def step_resident():
        for s in range(0, K, args.chunk):
            e = min(s + args.chunk, K)
            fg.render_composite_batch(cams_ngp[s:e], res, res, bg_image, bg_depth, out_u8=u8[: e - s], ngp_convention=True,
                                      rects_out=rects[: e - s], bg_u8_out=bg_u8)
            patches, _ = cv.preprocess(u8[: e - s], rot90=True, bg_u8=bg_u8, rects=rects[: e - s])
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            emb = cv.encode_patches(patches, e - s)
            ev[1].record()
            vit_events.append(ev)
            scores[s:e] = cv.score(emb, txt, n_goal=1)
        if world > 1:
            dist.all_gather_into_tensor(gathered, scores)

    pinned_poses = torch.from_numpy(valid_poses_ngp).pin_memory()
    host_scores = torch.empty(K, dtype=torch.float32).pin_memory()

    from dream2real_b200.clip_scoring import score_renders

    def step_e2e():
        """public API with HOST buffers, the way optimise_pose_grid drives it: poses in pinned host memory ->
        renderer.render (one call: background once, all candidates) -> score_renders -> scores back on the host."""
        vp = pinned_poses.numpy()
        out = rnd.render(vp, render_poses_ngp, [0], tm.depths[:1], tm.movable_masks, save=False, return_tensor=True)
        scores.copy_(score_renders(out, cv, txt, n_goal=1, bg_u8=rnd.last_bg_u8, rects=rnd.last_rects))
        del out
        if world > 1:
            dist.all_gather_into_tensor(gathered, scores)
        host_scores.copy_(scores, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    N.check(N.lib().d2r_profile_enable(local, 1))
    N.launch_count(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    vit_events.clear()
    ms = timed(step_resident, args.steps)
    vit_ms = sum(a.elapsed_time(b) for a, b in vit_events)
    clk = clocks.stop() if rank == 0 else None
    launches = N.launch_count()
    import ctypes as C
    mm, nl, ns, nt = C.c_float(), C.c_int(), C.c_ulonglong(), C.c_ulonglong()
    N.check(N.lib().d2r_profile_read(local, C.byref(mm), C.byref(nl), C.byref(ns), C.byref(nt)))
    N.check(N.lib().d2r_profile_enable(local, 0))
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    # algorithmic bytes of the march kernel: 512 B of hash-table reads per network
    # sample + 23 B per primary ray it owns (20 B cached background rgba+depth read, 3 B u8 written)
    rays = int(nt.value)          # primary rays (pixels of the candidates' screen rectangles) the kernel owned
    alg_bytes = int(ns.value) * 512 + rays * 23
    march_s = mm.value / 1e3
    achieved = alg_bytes / march_s / 1e9 if march_s > 0 else 0.0
    # DRAM traffic of one march launch from the committed ncu --set full capture of the same launch shape (profiles/)
    traffic = None
    try:
        summ = json.load(open(os.path.join(ROOT, "profiles", "march_ncu_summary.json")))
        if summ.get("resolution") == res and args.scene == summ.get("scene"):
            # captured on a 512-candidate launch; candidates are independent, so a launch of `chunk` candidates moves chunk/512 x as much
            traffic = summ["dram_bytes_per_launch"] * min(args.chunk, K) / summ["candidates_per_launch"]
    except Exception:
        pass
    total = K * world
    value = total * args.steps / (ms / 1e3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 (fp16 operands, fp32 accumulate; fp32 residual/softmax/LayerNorm)", "data": "synthetic",
        "config": {"workload": f"{args.scene} scene stand-in, {K} candidate poses per GPU, {res}x{res}, CLIP {args.clip} random-init, "
                               f"{world} x B200", "poses_per_gpu": K, "resolution": res, "clip": args.clip, "chunk": args.chunk,
                   "pose_grid": grid_res, "hash_table": f"2^{args.log2_hashmap}", "l2": "inputs larger than L2 (each chunk's u8 frames "
                   f"= {args.chunk * res * res * 3 / 1e6:.0f} MB)", "samples_per_candidate": int(ns.value) / max(1, K * args.steps),
                   "rays_marched_per_candidate": rays / max(1, K * args.steps)},
        "clocks": clk,
        "e2e": {"value": total * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(K * 12 * 4),
                "d2h_bytes_per_step": int(K * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "ray march (k_gather_round + k_mlp_round, all rounds of a launch; D2R_MARCH=fused: k_march_tc2)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes / max(1, nl.value), "peak_source": peak_src, "launches": int(nl.value),
                     "avg_launch_ms": mm.value / max(1, nl.value), "share_of_step": mm.value / ms,
                     "note": "algorithmic bytes (512 B of table reads per sample + 23 B per primary ray); the hash tables (~25 MB) are L2-resident, so DRAM traffic is "
                             "mostly the fp16 features the two march kernels hand over (DESIGN.md section 5); traffic = ncu dram bytes per launch"},
    }
    from dream2real_b200.clip import CLIP_CONFIGS
    c = CLIP_CONFIGS[args.clip]
    T = (c["image_size"] // c["patch_size"]) ** 2 + 1
    kp = (3 * c["patch_size"] ** 2 + 63) // 64 * 64
    d_, mlp_ = c["hidden"], c["mlp"]
    vit_flop = (2.0 * (T - 1) * d_ * kp + c["layers"] * (8.0 * T * d_ * d_ + 4.0 * T * d_ * mlp_ + 4.0 * T * T * d_) + 2.0 * d_ * c["proj"])
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    vit_tf = vit_flop * K * args.steps / (vit_ms / 1e3) / 1e12 if vit_ms > 0 else 0.0
    out["roofline_vit"] = {"kernel": "CLIP ViT forward (tcgen05 GEMMs + attention + LayerNorm)", "bound": "tensor", "achieved": vit_tf,
                           "peak": tf_peak, "unit": "TFLOP/s", "frac": vit_tf / tf_peak,
                           "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if "bf16_tflops_sustained" in peaks
                           else "fallback (B200_PROFILING.md)", "gflop_per_image": vit_flop / 1e9, "share_of_step": vit_ms / ms}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, scene_dir, n=args.cpu_sample)
    try:
        out["gpu_reference_context"] = {"what": "reference pyngp Shade+Depth pairs/s on B200 (fox snapshot, render only, no CLIP), tests/golden/pyngp_timing.json",
                                        **json.load(open(os.path.join(ROOT, "tests", "golden", "pyngp_timing.json")))}
    except Exception:
        pass
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_pipeline(args, scene_dir, n):
    """The reference algorithm restated on the CPU (oracle/): per candidate the NGP march of the movable object
    (colour + depth from one march, rays that miss the occupied box culled -- two result-preserving shortcuts the
    reference's two full-frame renders per candidate, combined_rendering.py:123-130, do not take), numpy composite,
    rot90, PIL preprocessing, HF CLIP fp32."""
    import torch

    from dream2real_b200 import ingp, synth
    from dream2real_b200.clip import make_hf_clip
    from oracle import ngp_oracle as O
    from oracle import post_oracle as PO
    torch.set_num_threads(os.cpu_count() or 1)
    scene = synth.make_scene(args.scene, scene_dir, log2_hashmap_size=args.log2_hashmap, seed=1234)
    fg = ingp.load_snapshot(os.path.join(scene_dir, "fg_base.ingp"))
    bg = ingp.load_snapshot(os.path.join(scene_dir, "bg_base.ingp"))
    res = args.res
    vs = O.view_setup(bg, 0, res, res)
    dirs = O.camera_plane_dirs(vs)
    fgb, _ = O.build_bitfield(fg.density_grid, fg.max_cascade)
    bgb, _ = O.build_bitfield(bg.density_grid, bg.max_cascade)
    box = O.occupied_box(fgb, fg.max_cascade)
    rp = PO.converter(scene["opt_cam_poses"][:1])
    # once per query (not timed, like the GPU arm): background render + depth, CLIP model, text
    bg_img = O.render(bg, bgb, vs, rp[0][:3], mode=O.SHADE, background_color=[0, 0, 0, 1], plane_dirs=dirs)
    bg_d = PO.background_depth(scene["depths"][0], scene["movable_masks"][0], (res, res))
    hf = make_hf_clip(args.clip, seed=1234, vocab_size=49408)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, 40000, (1 + len(NORM), 12), generator=g)
    ids[:, -1] = 2
    poses, _ = pose_grid(scene, args.poses)
    vp = PO.converter(poses)
    T1 = PO.converter(scene["fg_pose"][None])[0]
    R = hf.config.vision_config.image_size

    def step(idx):
        imgs = []
        for i in idx:
            cam = PO.convert_virtual_pose(T1, vp[i], rp[0])
            # one march for colour and depth and the occupied-box ray cull: both favour the CPU arm
            sh, dp = O.render(fg, fgb, vs, cam[:3], both=True, background_color=[0, 0, 0, 0], plane_dirs=dirs, cull_box=box)
            imgs.append(PO.composite(bg_img, bg_d, sh, dp[..., 0]))
        imgs = np.rot90(np.stack(imgs), k=1, axes=(1, 2))
        px = PO.clip_preprocess(imgs, R)
        logits = PO.clip_logits(hf, px, ids)
        return PO.normalise_scores(logits, 1)
    return step


def cpu_baseline(args, scene_dir, n):
    step = cpu_pipeline(args, scene_dir, n)
    stride = max(1, args.poses // n)
    idx = [(i * stride + stride // 2) % args.poses for i in range(n)]
    step(idx[:1])
    t0 = time.time()
    step(idx)
    dt = time.time() - t0
    return {"value": n / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} of the {args.poses} candidates (evenly spaced) at {args.res}x{args.res}, CLIP {args.clip} fp32: numpy oracle render "
                      f"(single thread; colour+depth in one march, occupied-box ray cull) + HF CLIP on {os.cpu_count()} torch threads; "
                      "the reference itself has no CPU render path (pyngp is CUDA-only)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    scene_dir = tempfile.mkdtemp(prefix="d2r_bench_ref_")
    n = args.cpu_sample
    step = cpu_pipeline(args, scene_dir, n)
    stride = max(1, args.poses // n)
    base = [(i * stride + stride // 2) for i in range(n)]
    for w in range(min(args.warmup, 1)):
        step([b % args.poses for b in base[:1]])
    t0 = time.time()
    for s in range(args.steps):
        step([(b + 7 * s) % args.poses for b in base])
    dt = time.time() - t0
    value = n * args.steps / dt
    sample = (f"each step = {n} of the {args.poses} candidates at {args.res}x{args.res} (bounded sample), CLIP {args.clip} fp32; numpy oracle "
              f"render (single thread) + HF CLIP on {os.cpu_count()} torch threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (numpy/torch CPU; fp16-emulated NGP network)", "data": "synthetic",
        "config": {"workload": f"{args.scene} scene stand-in, {args.poses} candidate poses, {args.res}x{args.res}, CLIP {args.clip} random-init, "
                               "CPU port of the reference path (the reference's renderer is CUDA-only)", "sample_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

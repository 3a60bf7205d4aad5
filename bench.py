#!/usr/bin/env python3
"""Benchmark of Dream2Real's imagination-and-scoring hot path (BASELINE.json metric:
candidate renders + CLIP scores per second at 800x800).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C3|C4|C5]
                    [--scene shopping] [--poses 4096] [--res 800] [--clip ViT-B/32] [--chunk 1024]

One "step" = one pass of the hot path over one batch of synthetic candidate poses: fused fg render +
depth-test composite -> rot90 + PIL-exact preprocess -> ViT forward on tcgen05 -> score.
--config picks a BASELINE.json configuration; the default is C2 = configs[1] (shopping scene stand-in,
4096 poses, 800x800, 1 x B200) for N = 1, 2, 4 -- every rank gets its own 4096 poses (weak scaling) -- and C4 (shelf scene,
6-DoF grid of 65 536 poses sharded 8192 per rank) for N = 8, whose line also carries the C2-per-rank weak-scaling figure
as `weak_scaling_same_workload`.  The ranks exchange scores with one NCCL all-gather per step.  Prints ONE JSON line (rank 0).
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
# BASELINE.json configs as concrete synthetic inputs; poses = per-GPU count at the named GPU count
CONFIGS = {
    "C1": dict(scene="shopping", grid=[8, 8, 1, 1, 1, 1], poses=64, res=400, gpus=1),
    "C2": dict(scene="shopping", grid=[64, 64, 1, 1, 1, 1], poses=4096, res=800, gpus=1),
    "C3": dict(scene="pool_triangle", grid=[128, 128, 1, 1, 1, 1], poses=16384, res=800, gpus=1),
    "C4": dict(scene="shelf", grid=[16, 4, 16, 4, 4, 4], poses=65536, res=800, gpus=8),
    "C5": dict(scene="synthetic8", grid=[64, 64, 1, 4, 4, 4], poses=262144, res=1024, gpus=8),
}
UNIT = "candidates/s"
GOAL = "an apple inside a blue and white bowl"          # reference lang/cache.json (shopping demo)
NORM = ["an apple and a blue and white bowl"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--scene", default=None)
    ap.add_argument("--poses", type=int, default=None, help="candidate poses per GPU")
    ap.add_argument("--res", type=int, default=None)
    ap.add_argument("--clip", default="ViT-B/32", choices=["ViT-B/32", "ViT-L/14-336"])
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--log2-hashmap", type=int, default=19)
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates per CPU-baseline sample / reference step (0 = one per host core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="N = 8 default run: skip the C2-per-rank weak-scaling figure")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    explicit = a.config is not None or a.scene is not None or a.poses is not None or a.res is not None
    if a.config is None:
        a.config = "C4" if (world == 8 and not explicit) else "C2"
    a.default_run = not explicit
    c = CONFIGS[a.config]
    a.scene = a.scene or c["scene"]
    a.res = a.res or c["res"]
    a.grid = c["grid"]
    if a.poses is None:
        # C2 is quoted per GPU (weak scaling); C4 / C5 are fixed-size grids sharded over the ranks that run them
        a.poses = c["poses"] if c["gpus"] == 1 else max(1, c["poses"] // max(world, 1))
    a.sharded = c["gpus"] > 1
    return a


def pose_grid(scene, n, grid=None, rank=0, world=1, sharded=False):
    """Candidate poses of this rank: sample_poses_grid(grid) for the scene type (x slowest ... z-rotation fastest).
    sharded: this rank's shard of the full grid (clip_scoring.shard_indices, strided); else the first n poses (repeated if the
    grid is smaller), which the caller offsets per rank."""
    import types

    import torch

    from dream2real_b200.clip_scoring import shard_indices
    from dream2real_b200.vision_3d.obj_pose_opt import sample_poses_grid
    if grid is None:
        g = int(np.ceil(np.sqrt(n)))
        grid = [g, g, 1, 1, 1, 1] if scene["scene_type"] != 1 else [max(2, int(round(n ** (1 / 3)))) for _ in range(3)] + [1, 1, 1]
    tm = types.SimpleNamespace(scene_model=types.SimpleNamespace(scene_centre=torch.tensor(scene["scene_centre"]), device=torch.device("cpu")))
    p = sample_poses_grid(tm, grid, scene_type=scene["scene_type"])
    if sharded:
        p = p[shard_indices(p.shape[0], world, rank)]      # strided: what optimise_pose_grid does (balanced ranks)
        n = min(n, p.shape[0]) if p.shape[0] else n
    reps = int(np.ceil(n / max(1, p.shape[0])))
    return p.repeat(reps, 1)[:n].reshape(-1, 4, 4).numpy().astype(np.float64), grid


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


def build_world(args, device, scene_dir, scene_name=None):
    """Scene, models, CLIP and cached background -- everything that is per query, not per candidate."""
    import torch

    from dream2real_b200 import synth
    from dream2real_b200.clip import ClipVision, make_hf_clip, text_embeds
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    scene = synth.make_scene(scene_name or args.scene, scene_dir, log2_hashmap_size=args.log2_hashmap, seed=1234)
    tm = synth.SyntheticTaskModel(scene, GOAL, NORM, device)
    rnd = renderer(scene_dir, tm, resolution=args.res, max_candidates_per_launch=args.chunk)
    hf = make_hf_clip(args.clip, seed=1234, vocab_size=49408)
    cv = ClipVision(hf, max_batch=args.chunk, device=device.index)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, 40000, (1 + len(NORM), 12), generator=g)
    ids[:, -1] = 2
    txt = text_embeds(hf, ids).to(device)
    return scene, tm, rnd, hf, cv, txt, accio2ngp


def march_sources_sha1():
    """content hash of the march kernel sources: ties a committed ncu traffic summary to the code it was captured on"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "dream2real_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith("d2r_march") or f in ("d2r_common.cuh", "d2r_gemm.cuh"):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def measure(args, world_ctx, scene_name, K, grid, sharded, steps, warmup, sample_clocks=True):
    """Device-resident value, e2e value, march roofline and ViT roofline of one workload on this rank set."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from dream2real_b200 import _native as N
    rank, world, local, device = world_ctx
    scene_dir = tempfile.mkdtemp(prefix=f"d2r_bench_r{rank}_")
    scene, tm, rnd, hf, cv, txt, accio2ngp = build_world(args, device, scene_dir, scene_name)
    res = args.res
    poses, grid_res = pose_grid(scene, K, grid, rank, world, sharded)
    K = poses.shape[0]
    if not sharded:
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

    def step_e2e():
        """public API with HOST buffers, the way optimise_pose_grid drives it: poses in pinned host memory -> renderer.iter_render
        (background once, then per chunk: fused render+composite -> preprocess -> ViT -> score) -> scores back on the host."""
        vp = pinned_poses.numpy()
        for _, s, e, frames, rc, bgu in rnd.iter_render(vp, render_poses_ngp, [0], tm.depths[:1], tm.movable_masks, save=False, chunk=args.chunk):
            emb = cv.encode_images(frames, rot90=True, bg_u8=bgu, rects=rc)
            scores[s:e] = cv.score(emb, txt, n_goal=1)
        if world > 1:
            dist.all_gather_into_tensor(gathered, scores)
        host_scores.copy_(scores, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(warmup, 3)):
        step_resident()
    N.check(N.lib().d2r_profile_enable(local, 1))
    N.launch_count(reset=True)
    clocks = ClockSampler(local)
    if rank == 0 and sample_clocks:
        clocks.start()
    vit_events.clear()
    ms = timed(step_resident, steps)
    vit_ms = sum(a.elapsed_time(b) for a, b in vit_events)
    clk = clocks.stop() if (rank == 0 and sample_clocks) else None
    launches = N.launch_count()
    mm, nl, ns, nt = C.c_float(), C.c_int(), C.c_ulonglong(), C.c_ulonglong()
    N.check(N.lib().d2r_profile_read(local, C.byref(mm), C.byref(nl), C.byref(ns), C.byref(nt)))
    N.check(N.lib().d2r_profile_enable(local, 0))
    step_e2e()
    ms_e2e = timed(step_e2e, steps)
    tot = torch.tensor([K], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tot)
    total = int(tot.item())
    # every rank's own time for the same timed region (each step ends with the all-gather, so these are near-equal; the
    # per-rank march time shows which GPU set the pace)
    mine = torch.tensor([float(mm.value) / max(1, steps)], device=device)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    march_ms_ranks = [round(float(t.item()), 2) for t in per_rank]
    return dict(scene=scene, scene_dir=scene_dir, K=K, total=total, grid=grid_res, ms=ms, ms_e2e=ms_e2e, vit_ms=vit_ms, clocks=clk,
                launches=int(launches), march_ms=float(mm.value), march_launches=int(nl.value), samples=int(ns.value), rays=int(nt.value),
                march_ms_ranks=march_ms_ranks)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = (rank, world, local, device)
    res, steps = args.res, args.steps
    m = measure(args, ctx, args.scene, args.poses, args.grid, args.sharded, steps, args.warmup)
    secondary = None
    if args.default_run and args.config == "C4" and not args.no_secondary:
        # the N = 1, 2, 4 lines are C2 per rank: the same workload at this N, so the scaling series stays comparable
        c2 = CONFIGS["C2"]
        secondary = measure(args, ctx, c2["scene"], c2["poses"], c2["grid"], False, steps, args.warmup, sample_clocks=False)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    K = m["K"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    # algorithmic bytes of the march kernel: 512 B of hash-table reads per network
    # sample + 23 B per primary ray it owns (20 B cached background rgba+depth read, 3 B u8 written)
    rays = m["rays"]          # primary rays (hit-list entries) the kernel owned
    alg_bytes = m["samples"] * 512 + rays * 23
    march_s = m["march_ms"] / 1e3
    achieved = alg_bytes / march_s / 1e9 if march_s > 0 else 0.0
    # DRAM / L2 traffic of one march launch: ncu capture of the same launch shape, committed under profiles/ together with the
    # hash of the kernel sources it was taken on -- a capture of other code is refused, not quoted
    traffic = l2_traffic = None
    traffic_note = "no ncu capture committed for this workload"
    try:
        summ = json.load(open(os.path.join(ROOT, "profiles", "march_ncu_summary.json")))
        if summ.get("source_sha1") != march_sources_sha1():
            traffic_note = "profiles/march_ncu_summary.json was captured on other kernel sources (source_sha1 mismatch): not quoted"
        elif summ.get("resolution") == res and m["scene"]["name"] == summ.get("scene"):
            # candidates are independent, so a launch of `chunk` candidates moves chunk / captured-candidates x as much
            f = min(args.chunk, K) / summ["candidates_per_launch"]
            traffic = summ["dram_bytes_per_launch"] * f
            l2_traffic = summ.get("lts_bytes_per_launch", 0) * f or None
            traffic_note = "ncu dram__bytes_read+write.sum / lts__t_bytes.sum over all march kernels of one launch, scaled to this launch size (profiles/march_ncu_summary.json)"
    except Exception:
        pass
    total = m["total"]
    value = total * steps / (m["ms"] / 1e3)
    launches_per_step = max(1, m["march_launches"] // max(1, steps))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": m["ms"] / steps, "higher_is_better": True, "scaling": "strong" if args.sharded else "weak", "vs_baseline": None,
        "dtype": "f16 (fp16 operands, fp32 accumulate; fp32 residual/softmax/LayerNorm)", "data": "synthetic",
        "config": {"workload": f"{args.config}: {m['scene']['name']} scene stand-in, "
                               + (f"{total} candidate poses (pose grid {m['grid']}) sharded over {world} x B200, {K} per GPU" if args.sharded
                                  else f"{K} candidate poses per GPU (pose grid {m['grid']}), {world} x B200")
                               + f", {res}x{res}, CLIP {args.clip} random-init",
                   "baseline_config": args.config, "poses_per_gpu": K, "poses_total": total, "resolution": res, "clip": args.clip, "chunk": args.chunk,
                   "pose_grid": m["grid"], "hash_table": f"2^{args.log2_hashmap}", "l2": "inputs larger than L2 (each chunk's u8 frames "
                   f"= {min(args.chunk, K) * res * res * 3 / 1e6:.0f} MB)", "samples_per_candidate": m["samples"] / max(1, K * steps),
                   "rays_marched_per_candidate": rays / max(1, K * steps)},
        "clocks": m["clocks"],
        "e2e": {"value": total * steps / (m["ms_e2e"] / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(K * 12 * 4),
                "d2h_bytes_per_step": int(K * 4), "ms_per_step": m["ms_e2e"] / steps,
                "path": "renderer.iter_render -> ClipVision.encode_images -> ClipVision.score per chunk (what optimise_pose_grid runs), pinned host poses in, pinned host scores out"},
        "gpu_launches": m["launches"],
        "roofline": {"kernel": "the ray march of one launch of `chunk` candidates: k_classify, 16 x (k_gather_round, k_mlp_round), k_march_ws (persistent tail), k_finish", "bound": "hbm",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "l2_traffic": l2_traffic,
                     "traffic_note": traffic_note, "algorithmic_bytes_per_launch": alg_bytes / max(1, m["march_launches"]), "peak_source": peak_src,
                     "launches": m["march_launches"], "launches_per_step": launches_per_step,
                     "avg_launch_ms": m["march_ms"] / max(1, m["march_launches"]), "share_of_step": m["march_ms"] / m["ms"],
                     "march_ms_per_step_by_rank": m["march_ms_ranks"],
                     "note": "algorithmic bytes (512 B of table reads per sample + 23 B per primary ray); the hash tables (~25 MB) are L2-resident, so "
                             "the table reads are L2 traffic, not DRAM traffic (SURVEY.md 8(d) caveat): `traffic` (DRAM) and `l2_traffic` (lts) are the "
                             "measured bytes per launch next to the algorithmic figure"},
    }
    from dream2real_b200.clip import CLIP_CONFIGS
    c = CLIP_CONFIGS[args.clip]
    T = (c["image_size"] // c["patch_size"]) ** 2 + 1
    kp = (3 * c["patch_size"] ** 2 + 63) // 64 * 64
    d_, mlp_ = c["hidden"], c["mlp"]
    vit_flop = (2.0 * (T - 1) * d_ * kp + c["layers"] * (8.0 * T * d_ * d_ + 4.0 * T * d_ * mlp_ + 4.0 * T * T * d_) + 2.0 * d_ * c["proj"])
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    vit_tf = vit_flop * K * steps / (m["vit_ms"] / 1e3) / 1e12 if m["vit_ms"] > 0 else 0.0
    out["roofline_vit"] = {"kernel": "CLIP ViT forward (tcgen05 GEMMs + attention + LayerNorm)", "bound": "tensor", "achieved": vit_tf,
                           "peak": tf_peak, "unit": "TFLOP/s", "frac": vit_tf / tf_peak,
                           "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if "bf16_tflops_sustained" in peaks
                           else "fallback (B200_PROFILING.md)", "gflop_per_image": vit_flop / 1e9, "share_of_step": m["vit_ms"] / m["ms"]}
    if secondary is not None:
        out["weak_scaling_same_workload"] = {
            "workload": f"C2 per rank: shopping scene stand-in, {secondary['K']} candidate poses per GPU, {res}x{res} -- the workload of the N = 1, 2, 4 lines",
            "value": secondary["total"] * steps / (secondary["ms"] / 1e3), "unit": UNIT, "ms_per_step": secondary["ms"] / steps,
            "e2e": secondary["total"] * steps / (secondary["ms_e2e"] / 1e3)}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
    try:
        t = json.load(open(os.path.join(ROOT, "tests", "golden", "pyngp_synth_timing.json")))
        r = t.get(str(res), t.get("800"))
        out["gpu_reference_context"] = {
            "what": "the reference's own GPU loop on the bench scene, run on a B200 by tests/golden/make_golden_synth.py (not part of this run): per candidate "
                    "two pyngp renders + NumPy composite (combined_rendering.py:117-155), then CLIPProcessor + HF CLIPModel fp32 in batches of 128 "
                    "(clip_scoring.py:168-185)", "scene": t.get("scene"), "resolution": res if str(res) in t else 800,
            "render_composite_ms_per_candidate": r["render_composite_ms_per_candidate"],
            "candidates_per_s": r.get(args.clip, {}).get("candidates_per_s")}
    except Exception:
        pass
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---- CPU arm: the reference algorithm restated on the CPU (oracle/), all host cores ---------------------------------------
_CPU = {}


def _cpu_init(scene_name, res, log2_hashmap, scene_dir):
    """per worker process: scene, snapshots, view table, background render (once per query, like the GPU arm)"""
    from dream2real_b200 import ingp, synth
    from oracle import ngp_oracle as O
    from oracle import post_oracle as PO
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    scene = synth.make_scene(scene_name, scene_dir, log2_hashmap_size=log2_hashmap, seed=1234) if not os.path.exists(os.path.join(scene_dir, "fg_base.ingp")) \
        else None
    _CPU["fg"] = ingp.load_snapshot(os.path.join(scene_dir, "fg_base.ingp"))
    _CPU["vs"] = O.view_setup(_CPU["fg"], 0, res, res)
    _CPU["dirs"] = O.camera_plane_dirs(_CPU["vs"])
    _CPU["fgb"], _ = O.build_bitfield(_CPU["fg"].density_grid, _CPU["fg"].max_cascade)
    _CPU["box"] = O.occupied_box(_CPU["fgb"], _CPU["fg"].max_cascade)
    del scene


def _cpu_render_bg_band(job):
    """rows b, b + nb, ... of the background render (once per query; fanned out only to shorten the set-up)"""
    from oracle import ngp_oracle as O
    cam, b, nb = job
    vs = _CPU["vs_bg"]
    mask = np.zeros((vs.H, vs.W), bool)
    mask[b::nb] = True
    return O.render(_CPU["bg"], _CPU["bgb"], vs, cam, mode=O.SHADE, background_color=[0, 0, 0, 1], plane_dirs=_CPU["dirs_bg"], pixel_mask=mask)[b::nb]


def _cpu_render_one(job):
    """one candidate: NGP march of the movable object (colour + depth from one march, rays that miss the occupied box culled --
    two result-preserving shortcuts the reference's two full-frame renders per candidate, combined_rendering.py:123-130, do not
    take) + the NumPy composite"""
    from oracle import ngp_oracle as O
    from oracle import post_oracle as PO
    cam, bg_img, bg_d = job
    sh, dp = O.render(_CPU["fg"], _CPU["fgb"], _CPU["vs"], cam[:3], both=True, background_color=[0, 0, 0, 0], plane_dirs=_CPU["dirs"], cull_box=_CPU["box"])
    return PO.composite(bg_img, bg_d, sh, dp[..., 0])


def _cpu_render_band(job):
    """rows b, b + nb, b + 2 nb, ... of one candidate (interleaved, so that every band sees its share of the object): the same
    march and composite as _cpu_render_one on those pixels"""
    from oracle import ngp_oracle as O
    from oracle import post_oracle as PO
    cam, b, nb = job
    vs = _CPU["vs"]
    mask = np.zeros((vs.H, vs.W), bool)
    mask[b::nb] = True
    sh, dp = O.render(_CPU["fg"], _CPU["fgb"], vs, cam[:3], both=True, background_color=[0, 0, 0, 0], plane_dirs=_CPU["dirs"], cull_box=_CPU["box"],
                      pixel_mask=mask)
    return PO.composite(_CPU["bg_img"][b::nb], _CPU["bg_d"][b::nb], sh[b::nb], dp[b::nb, :, 0])


def cpu_pipeline(scene_name, grid, n_poses, res, clip_name, log2_hashmap, procs, split=1):
    """Returns step(idx) -> scores for the candidates idx, the render fanned out over `procs` worker processes (one candidate each,
    or -- split > 1 -- every candidate's rows dealt out over `split` jobs, so that a step can be shorter than one candidate on
    one core), then rot90, PIL preprocessing and HF CLIP fp32 on all torch threads."""
    import multiprocessing as mp

    import torch

    from dream2real_b200 import ingp, synth
    from dream2real_b200.clip import make_hf_clip
    from oracle import ngp_oracle as O
    from oracle import post_oracle as PO
    scene_dir = tempfile.mkdtemp(prefix="d2r_bench_cpu_")
    scene = synth.make_scene(scene_name, scene_dir, log2_hashmap_size=log2_hashmap, seed=1234)
    bg = ingp.load_snapshot(os.path.join(scene_dir, "bg_base.ingp"))
    vs = O.view_setup(bg, 0, res, res)
    dirs = O.camera_plane_dirs(vs)
    bgb, _ = O.build_bitfield(bg.density_grid, bg.max_cascade)
    rp = PO.converter(scene["opt_cam_poses"][:1])
    # once per query (not timed, like the GPU arm): background render + depth, CLIP model, text
    bg_d = PO.background_depth(scene["depths"][0], scene["movable_masks"][0], (res, res))
    _CPU["bg"], _CPU["bgb"], _CPU["bg_d"], _CPU["vs_bg"], _CPU["dirs_bg"] = bg, bgb, bg_d, vs, dirs          # forked workers inherit these
    pool = mp.get_context("fork").Pool(procs, initializer=_cpu_init, initargs=(scene_name, res, log2_hashmap, scene_dir)) if procs > 1 else None
    if pool is None:
        _cpu_init(scene_name, res, log2_hashmap, scene_dir)
        bg_img = O.render(bg, bgb, vs, rp[0][:3], mode=O.SHADE, background_color=[0, 0, 0, 1], plane_dirs=dirs)
    else:
        bg_img = np.empty((res, res, 4), np.float32)
        for b, band in enumerate(pool.map(_cpu_render_bg_band, [(rp[0][:3], b, procs) for b in range(procs)], chunksize=1)):
            bg_img[b::procs] = band
        if split > 1:      # the band jobs composite their own rows: the workers need the finished background -> fork them again
            pool.close(); pool.join()
            _CPU["bg_img"] = bg_img
            pool = mp.get_context("fork").Pool(procs, initializer=_cpu_init, initargs=(scene_name, res, log2_hashmap, scene_dir))
    hf = make_hf_clip(clip_name, seed=1234, vocab_size=49408)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, 40000, (1 + len(NORM), 12), generator=g)
    ids[:, -1] = 2
    poses, _ = pose_grid(scene, n_poses, grid)
    vp = PO.converter(poses)
    T1 = PO.converter(scene["fg_pose"][None])[0]
    R = hf.config.vision_config.image_size
    _CPU["bg_img"] = bg_img
    torch.set_num_threads(os.cpu_count() or 1)

    def step(idx):
        if split > 1:
            jobs = [(PO.convert_virtual_pose(T1, vp[i], rp[0]), b, split) for i in idx for b in range(split)]
            bands = pool.map(_cpu_render_band, jobs, chunksize=1) if pool is not None else [_cpu_render_band(j) for j in jobs]
            imgs = np.empty((len(idx), res, res, 3), np.uint8)
            for j, (_, b, _n) in enumerate(jobs):
                imgs[j // split, b::split] = bands[j]
        else:
            jobs = [(PO.convert_virtual_pose(T1, vp[i], rp[0]), bg_img, bg_d) for i in idx]
            imgs = pool.map(_cpu_render_one, jobs, chunksize=1) if pool is not None else [_cpu_render_one(j) for j in jobs]
            imgs = np.stack(imgs)
        imgs = np.rot90(imgs, k=1, axes=(1, 2))
        px = PO.clip_preprocess(imgs, R)
        logits = PO.clip_logits(hf, px, ids)
        return PO.normalise_scores(logits, 1)
    step.close = (lambda: (pool.close(), pool.join())) if pool is not None else (lambda: None)
    return step


def cpu_baseline(args):
    """BASELINE.json configs[0] (C1), exactly: shopping scene, the 64 poses of the 8x8 grid, 400x400, the reference path on the CPU of this
    box -- oracle render fanned out over all host cores, HF CLIP fp32 on all torch threads."""
    cores = os.cpu_count() or 1
    c1 = CONFIGS["C1"]
    step = cpu_pipeline(c1["scene"], c1["grid"], c1["poses"], c1["res"], args.clip, args.log2_hashmap, cores)
    step(list(range(min(cores, c1["poses"]))))          # warm the worker processes
    t0 = time.time()
    step(list(range(c1["poses"])))
    dt = time.time() - t0
    step.close()
    return {"value": c1["poses"] / dt, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"C1 = BASELINE.json configs[0]: shopping scene, all {c1['poses']} poses of the 8x8 grid at {c1['res']}x{c1['res']}, CLIP {args.clip} fp32: "
                      f"numpy oracle render, one candidate per worker process on {cores} host cores (colour+depth in one march, occupied-box ray cull) + "
                      f"HF CLIP on {cores} torch threads; the reference itself has no CPU render path (pyngp is CUDA-only)"}


def run_reference(args):
    """The reference arm: the CPU restatement of the reference path on this arm's config (scene, resolution, CLIP) on all host
    cores.  Each step is a bounded sample of the workload: every candidate's rows are dealt out over all cores, and the number of
    candidates per step is chosen from the warm-up step's time so that the K timed steps end within REF_BUDGET_S (a few minutes
    whatever K the caller asks for); --cpu-sample fixes it instead."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    REF_BUDGET_S = 120.0
    cores = os.cpu_count() or 1
    total = int(np.prod(args.grid)) if args.sharded else args.poses
    step = cpu_pipeline(args.scene, args.grid, total, args.res, args.clip, args.log2_hashmap, cores, split=cores)
    t0 = time.time()
    step([total // 2])                                   # warm-up (worker start-up, first CLIP call) ...
    t0 = time.time()
    step([total // 3])                                   # ... and the time of one candidate
    t_one = time.time() - t0
    n = args.cpu_sample or int(max(1, min(cores, REF_BUDGET_S / max(args.steps, 1) / max(t_one, 1e-3))))
    stride = max(1, total // n)
    base = [(i * stride + stride // 2) for i in range(n)]
    t0 = time.time()
    for s in range(args.steps):
        step([(b + 7 * s) % total for b in base])
    dt = time.time() - t0
    step.close()
    value = n * args.steps / dt
    sample = (f"each step = {n} of the {total} candidates at {args.res}x{args.res} (bounded sample, evenly spaced; {n} chosen so that {args.steps} steps fit "
              f"{REF_BUDGET_S:.0f} s: one candidate takes {t_one:.1f} s), CLIP {args.clip} fp32; numpy oracle render, every candidate's rows dealt out over "
              f"{cores} worker processes on {cores} host cores + HF CLIP on {cores} torch threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps, "warmup": 2, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.sharded else "weak", "vs_baseline": None, "dtype": "f32 (numpy/torch CPU; fp16-emulated NGP network)", "data": "synthetic",
        "config": {"workload": f"{args.config}: {args.scene} scene stand-in, {total} candidate poses, {args.res}x{args.res}, CLIP {args.clip} random-init, "
                               "CPU port of the reference path (the reference's renderer is CUDA-only)", "baseline_config": args.config, "sample_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

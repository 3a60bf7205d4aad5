#!/usr/bin/env python3
"""Per-stage device timings of the hot path (CUDA events, warm, inputs larger than L2 where it matters).

    python tools/stage_bench.py [--clip ViT-B/32] [--chunk 512] [--res 800] [--out gpurun_out/stages.json]

Prints one JSON object: GEMM TFLOP/s for every dense contraction of the ViT at the bench batch, the whole
ViT forward (TFLOP/s against MEASURED_PEAKS.json bf16 sustained), preprocessing GB/s and the march stages.
This is a measuring tool for DESIGN.md / profiles/, not a bench contract line.
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


ITERS, WARM = None, None


def timed(fn, iters=10, warm=3):
    import torch
    iters, warm = ITERS or iters, WARM if WARM is not None else warm
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clip", default="ViT-B/32")
    ap.add_argument("--chunk", type=int, default=512)
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-march", action="store_true")
    ap.add_argument("--gemm-only", action="store_true")
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--warm", type=int, default=None)
    a = ap.parse_args()
    global ITERS, WARM
    ITERS, WARM = a.iters, a.warm
    import numpy as np
    import torch

    from dream2real_b200 import _native as N
    from dream2real_b200.clip import CLIP_CONFIGS, ClipVision, make_hf_clip
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    tf_burst = peaks.get("bf16_tflops", 1590.0)
    c = CLIP_CONFIGS[a.clip]
    T = (c["image_size"] // c["patch_size"]) ** 2 + 1
    npatch = T - 1
    d, mlp = c["hidden"], c["mlp"]
    kp = (3 * c["patch_size"] ** 2 + 63) // 64 * 64
    B = a.chunk
    out = {"clip": a.clip, "chunk": B, "peak_tflops_sustained": tf_peak, "peak_tflops_burst": tf_burst, "gemm": []}

    def gemm_case(name, M, Nn, K, mode):
        A = (torch.randn(M, K, device=dev) * 0.5).half()
        W = (torch.randn(Nn, K, device=dev) * 0.05).half()
        bias = torch.randn(Nn, device=dev)
        o = torch.empty(M, Nn, device=dev, dtype=torch.float32 if mode >= 2 else torch.float16)
        sp = N.stream_ptr()

        def run():
            N.check(N.lib().d2r_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, M, Nn, K, bias.data_ptr(), mode, o.data_ptr(), Nn, sp))
        ms = timed(run)
        Wt = W.t().contiguous()
        ms_t = timed(lambda: torch.matmul(A, Wt))
        fl = 2.0 * M * Nn * K
        out["gemm"].append({"name": name, "M": M, "N": Nn, "K": K, "mode": mode, "ms": ms, "tflops": fl / ms / 1e9,
                            "frac_sustained": fl / ms / 1e9 / tf_peak, "cublas_ms": ms_t, "cublas_tflops": fl / ms_t / 1e9})

    gemm_case("patch_embed", B * npatch, d, kp, 3)
    gemm_case("qkv", B * T, 3 * d, d, 0)
    gemm_case("attn_out", B * T, d, d, 2)
    gemm_case("fc1", B * T, mlp, d, 1)
    gemm_case("fc2", B * T, d, mlp, 2)

    if a.gemm_only:
        print(json.dumps(out))
        return
    hf = make_hf_clip(a.clip, seed=1234)
    cv = ClipVision(hf, max_batch=B, device=0)
    patches = (torch.randn(B * npatch, kp, device=dev) * 0.5).half()
    cv._patches.copy_(patches)
    ms = timed(lambda: cv.encode_patches(cv._patches, B), iters=5)
    L = c["layers"]
    fl = B * (2.0 * npatch * d * kp + L * (2.0 * T * d * 3 * d + 2.0 * T * d * d + 4.0 * T * d * mlp + 4.0 * T * T * d) + 2.0 * d * c["proj"])
    out["vit_forward"] = {"ms": ms, "images_per_s": B / ms * 1e3, "tflops": fl / ms / 1e9, "frac_sustained": fl / ms / 1e9 / tf_peak,
                          "gflop_per_image": fl / B / 1e9}

    u8 = torch.randint(0, 255, (B, a.res, a.res, 3), dtype=torch.uint8, device=dev)
    ms = timed(lambda: cv.preprocess(u8, rot90=True), iters=5)
    byts = B * (a.res * a.res * 3 + 2 * c["image_size"] * a.res * 3 + npatch * kp * 2)
    out["preprocess"] = {"ms": ms, "gbs": byts / ms / 1e6, "bytes": byts}

    if not a.skip_march:
        from dream2real_b200 import synth
        from dream2real_b200.reconstruction.combined_rendering import renderer
        from dream2real_b200.utils import accio2ngp
        sys.path.insert(0, ROOT)
        import bench
        scene_dir = tempfile.mkdtemp(prefix="d2r_stage_")
        scene = synth.make_scene("shopping", scene_dir, log2_hashmap_size=19, seed=1234)
        tm = synth.SyntheticTaskModel(scene, bench.GOAL, bench.NORM, dev)
        rnd = renderer(scene_dir, tm, resolution=a.res, max_candidates_per_launch=B)
        poses, _ = bench.pose_grid(scene, 4096)
        vp = accio2ngp.converter(poses)
        rp = accio2ngp.converter(scene["opt_cam_poses"][:1])
        fg = tm.movable_obj.vis_model
        bg_image, bg_depth = rnd.render_background(rp[0], 0, tm.depths[0], tm.movable_masks[0])
        fg.set_camera_to_training_view(0)
        T1 = accio2ngp.converter(scene["fg_pose"][None])[0]
        cams = T1 @ (np.linalg.inv(vp) @ T1) @ (np.linalg.inv(T1) @ rp[0])
        cams_ngp = fg.cams_to_ngp(cams[:, :3, :])
        frames = torch.empty((B, a.res, a.res, 3), dtype=torch.uint8, device=dev)
        sel = cams_ngp[1792:1792 + B]
        ms = timed(lambda: fg.render_composite_batch(sel, a.res, a.res, bg_image, bg_depth, out_u8=frames, ngp_convention=True), iters=5)
        out["render_composite_chunk"] = {"ms": ms, "candidates_per_s": B / ms * 1e3}

    s = json.dumps(out)
    print(s)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()

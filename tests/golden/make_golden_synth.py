#!/usr/bin/env python3
"""Golden vectors of the REAL reference renderer (pyngp) on the SYNTHETIC bench scenes, plus the
timing of the reference's own GPU loop on the bench scene (the "B-REF-GPU" arm of BASELINE.md section 3).

Test infrastructure: runs on a B200 box (`gpurun`), needs the reference's pyngp staged under the
git-ignored baseline/_ref/ngp (built from /root/reference/reconstruction/instant-ngp, see
make_golden_pyngp.py).  Nothing in the product, the -m gpu tests, smoke() or bench.py imports it.

For each scene (dream2real_b200.synth stand-ins, 2^19-entry tables, seed 1234 -- byte-identical
.ingp files are rebuilt by the tests from the same seed) it replays reference
reconstruction/combined_rendering.py:95-155 with pyngp:
    bg:  set_camera_to_training_view, background_color=[0,0,0,1], set_nerf_camera_matrix, Shade render
    fg:  per candidate: convert_virtual_pose, set_nerf_camera_matrix, Shade render, Depth render
    NumPy depth-test composite, un-premultiply, sRGB, u8, alpha threshold
and stores, cropped to the candidate's footprint to stay small: fg Shade / Depth / Cost (sample
count) float renders, the composited u8 frame, and a strided sample + a window of the bg render.

    python tests/golden/make_golden_synth.py golden [--scenes ...] [--res 336 800] [--ncand 2]
    python tests/golden/make_golden_synth.py time   [--scene shopping] [--res 336 800] [--n 48]
Outputs: gpurun_out/golden/synth_<scene>.npz, gpurun_out/golden/pyngp_synth_timing.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
OUT = os.path.join(ROOT, "gpurun_out", "golden")

from make_golden_pyngp import _import_pyngp, load_vm  # noqa: E402

# candidate grid per scene family: (sample_res, indices into the grid) -- poses that keep the object in view
CANDS = {
    "shopping": ([8, 8, 1, 1, 1, 1], [27, 36, 12]),
    "pool_triangle": ([8, 8, 1, 1, 1, 1], [27, 45, 18]),
    "shelf": ([4, 2, 4, 2, 2, 2], [37, 141, 70]),
}


def scene_inputs(name, scene_dir):
    from dream2real_b200 import synth
    from oracle import post_oracle as PO
    scene = synth.make_scene(name, scene_dir, log2_hashmap_size=19, seed=1234)
    res, idx = CANDS[name]
    grid = PO.sample_poses_grid(scene["scene_centre"], res, scene["scene_type"]).reshape(-1, 4, 4).numpy().astype(np.float64)
    poses = grid[idx]
    rp = PO.converter(scene["opt_cam_poses"][:1])
    vp = PO.converter(poses)
    T1 = PO.converter(scene["fg_pose"][None])[0]
    cams = np.stack([PO.convert_virtual_pose(T1, vp[i], rp[0]) for i in range(len(idx))])
    return scene, poses, rp, cams


def render(ngp, vm, cam, res, mode):
    vm.set_nerf_camera_matrix(np.matrix(cam)[:-1, :])
    vm.render_ground_truth = False
    vm.render_mode = getattr(ngp.RenderMode, mode)
    return np.asarray(vm.render(res, res, 1, True), dtype=np.float32)


def cmd_golden(args):
    from oracle import post_oracle as PO
    ngp = _import_pyngp()
    os.makedirs(OUT, exist_ok=True)
    for name in args.scenes:
        d = f"/tmp/d2r_gold_{name}"
        scene, poses, rp, cams = scene_inputs(name, d)
        fg = load_vm(ngp, os.path.join(d, "fg_base.ingp"))
        bg = load_vm(ngp, os.path.join(d, "bg_base.ingp"))
        out = {"poses": poses, "cams": cams, "render_pose": rp[0], "sample_res": np.array(CANDS[name][0]),
               "pose_idx": np.array(CANDS[name][1]), "res": np.array(args.res)}
        for res in args.res:
            # combined_rendering.py:95-113
            bg.set_camera_to_training_view(0)
            bg.background_color = [0.0, 0.0, 0.0, 1.0]
            bg_img = render(ngp, bg, rp[0], res, "Shade")
            bg_d = PO.background_depth(scene["depths"][0], scene["movable_masks"][0], (res, res))
            out[f"bg_strided_{res}"] = bg_img[::8, ::8].copy()
            w0 = res // 2 - res // 10
            out[f"bg_window_{res}"] = bg_img[w0:w0 + res // 5, w0:w0 + res // 5].copy()
            out[f"bg_window_origin_{res}"] = np.array([w0, w0])
            fg.set_camera_to_training_view(0)
            for i in range(min(args.ncand, len(cams))):
                sh = render(ngp, fg, cams[i], res, "Shade")        # :122-126
                dp = render(ngp, fg, cams[i], res, "Depth")        # :127-130
                cost = render(ngp, fg, cams[i], res, "Cost")
                u8 = PO.composite(bg_img, bg_d, sh, dp[..., 0])    # :133-155
                ys, xs = np.nonzero((sh[..., 3] > 0) | (dp[..., 3] > 0) | (cost[..., 0] > 0))
                assert ys.size, f"{name} candidate {i}: object not in view"
                m = 4
                y0, y1, x0, x1 = max(ys.min() - m, 0), min(ys.max() + m + 1, res), max(xs.min() - m, 0), min(xs.max() + m + 1, res)
                out[f"rect_{res}_{i}"] = np.array([y0, y1, x0, x1])
                out[f"shade_{res}_{i}"] = sh[y0:y1, x0:x1].copy()
                out[f"depth_{res}_{i}"] = dp[y0:y1, x0:x1, 0].copy()
                out[f"depth_a_{res}_{i}"] = dp[y0:y1, x0:x1, 3].copy()
                out[f"cost_{res}_{i}"] = (cost[y0:y1, x0:x1, 0] * 128).astype(np.float32)
                out[f"u8_{res}_{i}"] = u8[y0:y1, x0:x1].copy()
                # what the crop leaves out: the composite outside it is the pure background composite
                outside = np.ones((res, res), bool)
                outside[y0:y1, x0:x1] = False
                empty = PO.composite(bg_img, bg_d, np.zeros_like(sh), np.zeros_like(dp[..., 0]))
                assert np.array_equal(u8[outside], empty[outside])
                out[f"u8_bg_crc_{res}_{i}"] = np.array([int(u8[outside].astype(np.uint64).sum())])
                print(f"[{name}] res {res} cand {i}: rect {y0}:{y1} x {x0}:{x1}, alpha>0 {(sh[..., 3] > 0).sum()} px, "
                      f"samples {cost[..., 0].sum() * 128:.0f}, depth max {dp[..., 0].max():.3f}", flush=True)
        path = os.path.join(OUT, f"synth_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)
        del fg, bg
        ngp.free_temporary_memory()
    print("GOLDEN SYNTH DONE", flush=True)


def cmd_time(args):
    """The reference's GPU loop on the bench scene: per candidate two pyngp renders (Shade, Depth) at res x res with the
    results copied to the host (python_api.cu:123-201), the NumPy composite (combined_rendering.py:133-155), then
    CLIPProcessor + CLIPModel fp32 in batches of 128 (clip_scoring.py:150-185) on the same GPU."""
    import torch
    from transformers import CLIPImageProcessor

    from dream2real_b200.clip import make_hf_clip
    from oracle import post_oracle as PO
    ngp = _import_pyngp()
    os.makedirs(OUT, exist_ok=True)
    d = f"/tmp/d2r_gold_{args.scene}"
    from dream2real_b200 import synth
    scene = synth.make_scene(args.scene, d, log2_hashmap_size=19, seed=1234)
    g = int(np.ceil(np.sqrt(args.n)))
    grid = PO.sample_poses_grid(scene["scene_centre"], [g, g, 1, 1, 1, 1], scene["scene_type"]).reshape(-1, 4, 4).numpy().astype(np.float64)[:args.n]
    rp = PO.converter(scene["opt_cam_poses"][:1])
    vp = PO.converter(grid)
    T1 = PO.converter(scene["fg_pose"][None])[0]
    fg = load_vm(ngp, os.path.join(d, "fg_base.ingp"))
    bg = load_vm(ngp, os.path.join(d, "bg_base.ingp"))
    result = {"scene": args.scene, "n_candidates": args.n, "gpu": torch.cuda.get_device_name(0)}
    for res in args.res:
        bg.set_camera_to_training_view(0)
        bg.background_color = [0.0, 0.0, 0.0, 1.0]
        bg_img = render(ngp, bg, rp[0], res, "Shade")
        bg_d = PO.background_depth(scene["depths"][0], scene["movable_masks"][0], (res, res))
        fg.set_camera_to_training_view(0)

        def one(i):
            cam = PO.convert_virtual_pose(T1, vp[i], rp[0])
            sh = render(ngp, fg, cam, res, "Shade")
            dp = render(ngp, fg, cam, res, "Depth")
            return PO.composite(bg_img, bg_d, sh, dp[..., 0])
        one(0)
        torch.cuda.synchronize()
        t0 = time.time()
        imgs = [one(i) for i in range(args.n)]
        torch.cuda.synchronize()
        t_render = (time.time() - t0) / args.n
        imgs = [np.rot90(im, k=1, axes=(0, 1)) for im in imgs]                # clip_scoring.py:145-147
        entry = {"render_composite_ms_per_candidate": t_render * 1e3, "render_candidates_per_s": 1.0 / t_render}
        for clip_name in ("ViT-B/32", "ViT-L/14-336"):
            hf = make_hf_clip(clip_name, seed=1234, vocab_size=49408).cuda().eval()
            R = hf.config.vision_config.image_size
            proc = CLIPImageProcessor(size={"shortest_edge": R}, crop_size={"height": R, "width": R})
            ids = torch.randint(3, 40000, (2, 12), generator=torch.Generator().manual_seed(1234))
            ids[:, -1] = 2
            ids = ids.cuda()

            def clip_pass():
                with torch.no_grad():
                    for s in range(0, len(imgs), 128):                         # clip_scoring.py:168-185
                        px = proc(images=[np.ascontiguousarray(im) for im in imgs[s:s + 128]], return_tensors="pt")["pixel_values"].cuda()
                        hf(pixel_values=px, input_ids=ids).logits_per_image.cpu()
            clip_pass()
            torch.cuda.synchronize()
            t0 = time.time()
            clip_pass()
            torch.cuda.synchronize()
            t_clip = (time.time() - t0) / len(imgs)
            entry[clip_name] = {"clip_ms_per_candidate": t_clip * 1e3, "candidates_per_s": 1.0 / (t_render + t_clip)}
            print(f"REFERENCE GPU loop @{res}x{res} {clip_name}: render+composite {t_render * 1e3:.2f} ms, CLIP fp32 {t_clip * 1e3:.2f} ms "
                  f"-> {1.0 / (t_render + t_clip):.2f} candidates/s", flush=True)
            del hf
            torch.cuda.empty_cache()
        result[str(res)] = entry
    json.dump(result, open(os.path.join(OUT, "pyngp_synth_timing.json"), "w"), indent=1)
    print("TIMING DONE", json.dumps(result), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("golden")
    p.add_argument("--scenes", nargs="+", default=["shopping", "pool_triangle", "shelf"])
    p.add_argument("--res", type=int, nargs="+", default=[336, 800])
    p.add_argument("--ncand", type=int, default=2)
    p.set_defaults(fn=cmd_golden)
    p = sub.add_parser("time")
    p.add_argument("--scene", default="shopping")
    p.add_argument("--res", type=int, nargs="+", default=[336, 800])
    p.add_argument("--n", type=int, default=48)
    p.set_defaults(fn=cmd_time)
    a = ap.parse_args()
    a.fn(a)

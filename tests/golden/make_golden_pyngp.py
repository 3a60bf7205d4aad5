#!/usr/bin/env python3
"""Golden-vector generator: runs the REAL reference renderer (pyngp, built from
/root/reference/reconstruction/instant-ngp by the survey round and staged under the
git-ignored baseline/_ref/ngp) on a B200 box and dumps its outputs.

This script is test infrastructure.  It is the only place the reference binary is ever
executed; nothing in the product, the -m gpu tests, smoke() or bench.py imports it.

It replays the exact call sequence of reference reconstruction/combined_rendering.py:98-130
(set_camera_to_training_view -> background_color -> set_nerf_camera_matrix -> render_mode
Shade/Depth -> render(w, h, 1, True)) and of ngp_visual_model.py:21-29 (Testbed(Nerf) +
load_snapshot).

Usage (on the GPU box, from the repo root):
    python tests/golden/make_golden_pyngp.py train  [--steps N]
    python tests/golden/make_golden_pyngp.py render SNAP.ingp CAMS.npy OUT.npz --res 64 128 [--view 0] [--bg 0 0 0 0]
    python tests/golden/make_golden_pyngp.py time   SNAP.ingp CAMS.npy --res 336 800
Outputs go under gpurun_out/golden/.
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "baseline", "_ref", "ngp")
OUT = os.path.join(ROOT, "gpurun_out", "golden")


def _import_pyngp():
    import torch  # noqa: F401  (loads libcudart.so.12 so that pyngp's DT_NEEDED resolves)
    sys.path.insert(0, REF)
    import pyngp as ngp
    return ngp


SMALL_CFG_PATCH = {"encoding": {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 4,
                                "log2_hashmap_size": 14, "base_resolution": 16}}


def describe(vm):
    md = vm.nerf.training.dataset.metadata[0]
    return {
        "aabb": [list(map(float, vm.aabb.min)), list(map(float, vm.aabb.max))],
        "render_aabb": [list(map(float, vm.render_aabb.min)), list(map(float, vm.render_aabb.max))],
        "background_color": list(map(float, vm.background_color)),
        "snap_to_pixel_centers": bool(vm.snap_to_pixel_centers),
        "exposure": float(vm.exposure),
        "min_transmittance": float(vm.nerf.render_min_transmittance),
        "cone_angle_constant": float(vm.nerf.cone_angle_constant),
        "fov_axis": int(vm.fov_axis),
        "screen_center": list(map(float, vm.screen_center)),
        "zoom": float(vm.zoom),
        "dataset_scale": float(vm.nerf.training.dataset.scale),
        "dataset_offset": list(map(float, vm.nerf.training.dataset.offset)),
        "dataset_aabb_scale": int(vm.nerf.training.dataset.aabb_scale),
        "n_images": int(vm.nerf.training.dataset.n_images),
        "view0_focal": list(map(float, md.focal_length)),
        "view0_pp": list(map(float, md.principal_point)),
        "view0_res": list(map(int, md.resolution)),
        "view0_lens_mode": str(md.lens.mode),
        "view0_lens_params": list(map(float, md.lens.params)),
        "n_params": int(vm.n_params()),
        "n_encoding_params": int(vm.n_encoding_params()),
    }


def load_vm(ngp, snap):
    # reference ngp_visual_model.py:24-28
    vm = ngp.Testbed(ngp.TestbedMode.Nerf)
    vm.load_snapshot(snap)
    return vm


def render_all(ngp, vm, cams, res, view, bg, modes=("Shade", "Depth", "Cost")):
    """reference combined_rendering.py:98-130 call order."""
    out = {}
    vm.set_camera_to_training_view(view)
    if bg is not None:
        vm.background_color = list(bg)
    for m in modes:
        imgs = []
        for c in cams:
            vm.set_nerf_camera_matrix(np.matrix(c)[:3, :])
            vm.render_ground_truth = False
            vm.render_mode = getattr(ngp.RenderMode, m)
            imgs.append(np.asarray(vm.render(res, res, 1, True), dtype=np.float32))
        out[m] = np.stack(imgs)
    return out


def cmd_train(args):
    ngp = _import_pyngp()
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(REF, "data", "fox")
    dst = "/tmp/fox_a2"  # dataset copy stays off gpurun_out (64 MiB cap)
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    tf = json.load(open(os.path.join(dst, "transforms.json")))
    tf["aabb_scale"] = 2  # what utils/accio2ngp.py:60 writes for Dream2Real scans
    tf["frames"] = [f for f in tf["frames"] if os.path.exists(os.path.join(dst, f["file_path"]))]
    json.dump(tf, open(os.path.join(dst, "transforms.json"), "w"), indent=1)
    json.dump(tf, open(os.path.join(OUT, "fox_a2_transforms.json"), "w"), indent=1)
    base_cfg = json.load(open(os.path.join(REF, "configs", "nerf", "base.json")))

    # camera set: training frame 0 plus perturbed copies and two other frames
    rng = np.random.default_rng(0)
    base = np.array(tf["frames"][0]["transform_matrix"], dtype=np.float64)
    cams = [base.copy()]
    for i in range(3):
        m = base.copy()
        m[:3, 3] += rng.normal(0, 0.25, 3)
        cams.append(m)
    cams.append(np.array(tf["frames"][7]["transform_matrix"], dtype=np.float64))
    cams.append(np.array(tf["frames"][23]["transform_matrix"], dtype=np.float64))
    cams = np.stack(cams)
    np.save(os.path.join(OUT, "fox_cams.npy"), cams)

    for tag, patch in (("small", SMALL_CFG_PATCH), ("full", {})):
        cfg = dict(base_cfg)
        cfg.update(patch)
        cfg_path = os.path.join(OUT, f"cfg_{tag}.json")
        json.dump(cfg, open(cfg_path, "w"))
        # reference reconstruction/train_ngp.py:42-92 (build_vis_model) call order
        tb = ngp.Testbed()
        tb.root_dir = REF
        tb.load_file(os.path.join(dst, "transforms.json"))
        tb.reload_network_from_file(cfg_path)
        tb.shall_train = True
        tb.nerf.render_with_lens_distortion = True
        tb.background_color = [0.0, 0.0, 0.0, 0.0]
        tb.nerf.training.near_distance = 0.1
        t0 = time.time()
        while tb.frame():
            if tb.training_step >= args.steps:
                break
        print(f"[{tag}] trained {tb.training_step} steps in {time.time()-t0:.1f}s loss={tb.loss:.5f}", flush=True)
        snap = os.path.join(OUT, f"fox_a2_{tag}.ingp")
        tb.save_snapshot(snap, False)
        print(f"[{tag}] snapshot bytes {os.path.getsize(snap)} n_params {tb.n_params()}", flush=True)
        del tb
        ngp.free_temporary_memory()

        vm = load_vm(ngp, snap)
        info = describe(vm)
        json.dump(info, open(os.path.join(OUT, f"fox_a2_{tag}_info.json"), "w"), indent=1)
        print(tag, json.dumps(info), flush=True)
        for res in ((48, 96, 336) if tag == "small" else (96, 336)):
            for bgname, bg in (("bg0", [0, 0, 0, 0]), ("bg1", [0, 0, 0, 1])):
                if bgname == "bg1" and res != 96:
                    continue
                cc = cams[:2] if res >= 336 else cams
                r = render_all(ngp, vm, cc, res, 0, bg)
                np.savez_compressed(os.path.join(OUT, f"fox_a2_{tag}_{res}_{bgname}.npz"),
                                    cams=cc, **r)
                a = r["Shade"][..., 3]
                print(f"[{tag}] res {res} {bgname}: alpha>0 {float((a>0).mean()):.3f} alpha>0.5 {float((a>0.5).mean()):.3f} "
                      f"depth max {float(r['Depth'][...,0].max()):.3f} cost max {float(r['Cost'][...,0].max()*128):.0f}", flush=True)
        if tag == "full":
            time_renders(ngp, vm, cams, (336, 800))
        del vm
        ngp.free_temporary_memory()
    print("GOLDEN TRAIN DONE", flush=True)


def time_renders(ngp, vm, cams, ress):
    import torch
    vm.set_camera_to_training_view(0)
    res_out = {}
    for res in ress:
        def pair(c):
            vm.set_nerf_camera_matrix(np.matrix(c)[:3, :])
            vm.render_ground_truth = False
            vm.render_mode = ngp.RenderMode.Shade
            a = vm.render(res, res, 1, True)
            vm.render_mode = ngp.RenderMode.Depth
            b = vm.render(res, res, 1, True)
            return a, b
        pair(cams[0])
        torch.cuda.synchronize()
        t0 = time.time()
        n = 0
        for _ in range(4):
            for c in cams:
                pair(c)
                n += 1
        dt = time.time() - t0
        res_out[str(res)] = n / dt
        print(f"REFERENCE pyngp Shade+Depth pairs @ {res}x{res}: {n/dt:.2f} candidates/s ({1000*dt/n:.2f} ms each)", flush=True)
    json.dump(res_out, open(os.path.join(OUT, "pyngp_timing.json"), "w"))


def cmd_render(args):
    ngp = _import_pyngp()
    os.makedirs(OUT, exist_ok=True)
    vm = load_vm(ngp, args.snapshot)
    info = describe(vm)
    print(json.dumps(info), flush=True)
    cams = np.load(args.cams)
    res_all = {}
    for res in args.res:
        r = render_all(ngp, vm, cams, res, args.view, args.bg)
        for k, v in r.items():
            res_all[f"{k}_{res}"] = v
        a = r["Shade"][..., 3]
        print(f"res {res}: alpha>0 {float((a>0).mean()):.4f} alpha>0.5 {float((a>0.5).mean()):.4f} "
              f"cost max {float(r['Cost'][...,0].max()*128):.0f}", flush=True)
    np.savez_compressed(args.out, cams=cams, info=json.dumps(info), **res_all)
    print("GOLDEN RENDER DONE", args.out, flush=True)


def cmd_time(args):
    ngp = _import_pyngp()
    vm = load_vm(ngp, args.snapshot)
    time_renders(ngp, vm, np.load(args.cams), args.res)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("train"); p.add_argument("--steps", type=int, default=2500); p.set_defaults(fn=cmd_train)
    p = sub.add_parser("render"); p.add_argument("snapshot"); p.add_argument("cams"); p.add_argument("out")
    p.add_argument("--res", type=int, nargs="+", default=[64]); p.add_argument("--view", type=int, default=0)
    p.add_argument("--bg", type=float, nargs=4, default=None); p.set_defaults(fn=cmd_render)
    p = sub.add_parser("time"); p.add_argument("snapshot"); p.add_argument("cams")
    p.add_argument("--res", type=int, nargs="+", default=[336, 800]); p.set_defaults(fn=cmd_time)
    a = ap.parse_args()
    a.fn(a)

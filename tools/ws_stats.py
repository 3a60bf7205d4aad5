#!/usr/bin/env python3
"""Where k_march_ws spends its cycles, per warp role (run under gpurun, 1 GPU):
    python tools/ws_stats.py [--poses 1024] [--res 800] [--scene shopping] [--chunk 1024]
Renders one chunk of candidates with profiling enabled and prints the role statistics the kernel accumulates
(d2r_profile_read_stats): share of the gather warps' time spent waiting for their MLP round / walking / gathering,
share of the epilogue warps' time spent waiting for an MMA, cycles per epilogue item, MMA-thread issue share."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=1024)
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--scene", default="shopping")
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch

    import bench
    from dream2real_b200 import _native as N
    from dream2real_b200 import synth
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    dev = torch.device("cuda", 0)
    d = tempfile.mkdtemp(prefix="d2r_ws_stats_")
    scene = synth.make_scene(a.scene, d, log2_hashmap_size=19, seed=1234)
    tm = synth.SyntheticTaskModel(scene, "g", None, dev)
    rnd = renderer(d, tm, resolution=a.res, max_candidates_per_launch=a.chunk)
    poses, _ = bench.pose_grid(scene, a.poses, bench.CONFIGS["C2"]["grid"] if a.scene == "shopping" else None)
    vp = accio2ngp.converter(poses)
    rp = accio2ngp.converter(scene["opt_cam_poses"][:1])
    fg = tm.movable_obj.vis_model
    bg_image, bg_depth = rnd.render_background(rp[0], 0, tm.depths[0], tm.movable_masks[0])
    fg.set_camera_to_training_view(0)
    T1 = accio2ngp.converter(scene["fg_pose"][None])[0]
    cams = fg.cams_to_ngp((T1 @ (np.linalg.inv(vp) @ T1) @ (np.linalg.inv(T1) @ rp[0]))[:, :3, :])
    u8 = torch.empty((a.chunk, a.res, a.res, 3), dtype=torch.uint8, device=dev)

    def run():
        for s in range(0, a.poses, a.chunk):
            e = min(s + a.chunk, a.poses)
            fg.render_composite_batch(cams[s:e], a.res, a.res, bg_image, bg_depth, out_u8=u8[: e - s], ngp_convention=True)
    run()
    torch.cuda.synchronize()
    N.check(N.lib().d2r_profile_enable(0, 1))
    for _ in range(a.reps):
        run()
    mm, nl, ns, nt = C.c_float(), C.c_int(), C.c_ulonglong(), C.c_ulonglong()
    N.check(N.lib().d2r_profile_read(0, C.byref(mm), C.byref(nl), C.byref(ns), C.byref(nt)))
    st = (C.c_ulonglong * 16)()
    N.check(N.lib().d2r_profile_read_stats(0, st))
    N.check(N.lib().d2r_profile_enable(0, 0))
    st = [int(x) for x in st]
    g_tot, g_wait, g_walk, g_enc, g_refill, e_tot, e_wait, e_items, m_tot, m_issue, g_rounds = st[2], st[3], st[4], st[5], st[6], st[7], st[8], st[9], st[10], st[11], st[12]
    out = {
        "march_ms_per_launch": mm.value / max(1, nl.value), "launches": nl.value, "samples": st[0], "rays": st[1],
        "gsamples_per_s": st[0] / (mm.value / 1e3) / 1e9 if mm.value else 0,
        "gather": {"wait_for_mlp": g_wait / max(1, g_tot), "walk": g_walk / max(1, g_tot), "hash_gather": g_enc / max(1, g_tot), "refill": g_refill / max(1, g_tot),
                   "cycles_per_round": g_tot / max(1, g_rounds), "samples_per_round_per_warp": st[0] / max(1, g_rounds)},
        "epilogue": {"wait_for_mma": e_wait / max(1, e_tot), "cycles_per_item_busy": (e_tot - e_wait) / max(1, e_items), "cycles_per_item_total": e_tot / max(1, e_items)},
        "mma": {"issuing": m_issue / max(1, m_tot)},
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""-m gpu: the other BASELINE.json scene families at sizes the oracle finishes in seconds -- 6-DoF (shelf, scene_type 1:
rotated candidate poses -> rotated virtual cameras), the pool_triangle bounds, a second view -- plus the
size-independent properties the path offers at any K: candidates are independent (a batch equals its parts, in any
chunking), a candidate at the object's initial pose reproduces the plain fg render, and candidates whose object leaves
the frame return the background."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_frames(scene, d, res, poses_nerf, view=0):
    import os

    from dream2real_b200 import ingp
    from oracle import ngp_oracle as O
    from oracle import post_oracle as PO
    fg, bg = ingp.load_snapshot(os.path.join(d, "fg_base.ingp")), ingp.load_snapshot(os.path.join(d, "bg_base.ingp"))
    vs = O.view_setup(bg, view, res, res)
    dirs = O.camera_plane_dirs(vs)
    fgb, _ = O.build_bitfield(fg.density_grid, fg.max_cascade)
    bgb, _ = O.build_bitfield(bg.density_grid, bg.max_cascade)
    vp = PO.converter(poses_nerf)
    rp = PO.converter(scene["opt_cam_poses"][view:view + 1])
    bg_img = O.render(bg, bgb, vs, rp[0][:3], mode=O.SHADE, background_color=[0, 0, 0, 1], plane_dirs=dirs)
    bg_d = PO.background_depth(scene["depths"][view], scene["movable_masks"][view], (res, res))
    T1 = PO.converter(scene["fg_pose"][None])[0]
    out = []
    for i in range(vp.shape[0]):
        cam = PO.convert_virtual_pose(T1, vp[i], rp[0])
        sh, dp = O.render(fg, fgb, vs, cam[:3], both=True, background_color=[0, 0, 0, 0], plane_dirs=dirs)
        out.append(PO.composite(bg_img, bg_d, sh, dp[..., 0]))
    return np.stack(out)


def _grid_poses(scene, sample_res, idx):
    from oracle import post_oracle as PO
    p = PO.sample_poses_grid(scene["scene_centre"], sample_res, scene["scene_type"])
    return p[idx].numpy().reshape(-1, 4, 4).astype(np.float64)


@pytest.mark.parametrize("name,sample_res,pick", [
    ("shelf", [3, 2, 3, 2, 2, 2], [0, 9, 35, 58, 77, 93]),        # 6-DoF: Euler ranges [-pi, pi/2] (obj_pose_opt.py:23-29)
    ("pool_triangle", [5, 5, 1, 1, 1, 1], [0, 7, 12, 18, 24]),       # scene_type 0 bounds (obj_pose_opt.py:16-22)
])
def test_scene_families_match_oracle(tmp_path, name, sample_res, pick):
    import torch
    from dream2real_b200 import synth
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    d = str(tmp_path)
    res = 64
    scene = synth.make_scene(name, d, log2_hashmap_size=13, seed=7)
    tm = synth.SyntheticTaskModel(scene, "g", None, torch.device("cuda"))
    poses = _grid_poses(scene, sample_res, pick)
    r = renderer(d, tm, resolution=res)
    got = r.render(accio2ngp.converter(poses), accio2ngp.converter(scene["opt_cam_poses"][:1]), [0], tm.depths[:1], tm.movable_masks,
                   save=False, return_tensor=True).cpu().numpy()
    ref = _oracle_frames(scene, d, res, poses)
    diff = np.abs(got.astype(int) - ref.astype(int))
    print(name, "u8 diff >1 LSB:", float((diff > 1).mean()), "max", diff.max(), "fg pixels changed vs bg:", int((ref != ref[:1]).any(-1).sum()))
    assert (diff > 1).mean() < 5e-3      # tolerance of tests/test_e2e_gpu.py: 1 LSB except single-sample flips at cell boundaries


def test_second_view_and_multi_view_order(tmp_path):
    """render_cam_pose_idx with two views returns view-major K*L frames (combined_rendering.py:95-163)."""
    import torch
    from dream2real_b200 import synth
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    d = str(tmp_path)
    res = 48
    scene = synth.make_scene("shopping", d, log2_hashmap_size=12, seed=3)
    tm = synth.SyntheticTaskModel(scene, "g", None, torch.device("cuda"))
    poses = _grid_poses(scene, [3, 3, 1, 1, 1, 1], [1, 4, 8])
    r = renderer(d, tm, resolution=res)
    r.fix_mask_index = True
    both = r.render(accio2ngp.converter(poses), accio2ngp.converter(scene["opt_cam_poses"][:2]), [0, 1], tm.depths[:2], tm.movable_masks,
                    save=False, return_tensor=True).cpu().numpy()
    assert both.shape == (6, res, res, 3)
    for view in (0, 1):
        ref = _oracle_frames(scene, d, res, poses, view=view)
        diff = np.abs(both[3 * view:3 * view + 3].astype(int) - ref.astype(int))
        assert (diff > 1).mean() < 5e-3, (view, float((diff > 1).mean()))


def test_candidates_are_independent_at_full_resolution(tmp_path):
    """Size-independent properties at the bench resolution (800x800, 2^19-entry tables): any chunking of a batch (different
    occupancy of the persistent kernel's ray slots, different hit-list order) gives bit-identical frames; the initial pose reproduces the plain composite of an un-moved object; an object moved out
    of the frame leaves exactly the background."""
    import torch
    from dream2real_b200 import synth
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    d = str(tmp_path)
    res = 800
    scene = synth.make_scene("shopping", d, log2_hashmap_size=19, seed=1234)
    tm = synth.SyntheticTaskModel(scene, "g", None, torch.device("cuda"))
    poses = _grid_poses(scene, [64, 64, 1, 1, 1, 1], list(range(0, 4096, 97)))      # 43 candidates across the grid
    far = scene["fg_pose"].copy()[None].astype(np.float64)
    far[0, :3, 3] += [30.0, 30.0, 0.0]                                             # far outside the view frustum
    poses = np.concatenate([poses, scene["fg_pose"][None].astype(np.float64), far])
    vp = accio2ngp.converter(poses)
    rp = accio2ngp.converter(scene["opt_cam_poses"][:1])
    outs = []
    for chunk in (64, 7, 1):
        r = renderer(d, tm, resolution=res, max_candidates_per_launch=chunk)
        sel = slice(None) if chunk != 1 else slice(40, 45)
        outs.append((sel, r.render(vp[sel], rp, [0], tm.depths[:1], tm.movable_masks, save=False, return_tensor=True)))
    full = outs[0][1]
    assert torch.equal(full, outs[1][1])
    assert torch.equal(full[40:45], outs[2][1])
    # background only: the far candidate equals a frame no ray touched -> every pixel equals the fill (composited background)
    bg_only = full[-1]
    r = renderer(d, tm, resolution=res)
    bg_image, bg_depth = r.render_background(rp[0], 0, tm.depths[0], tm.movable_masks[0])
    assert bg_only.shape == (res, res, 3) and bg_image.shape == (res, res, 4)
    # every candidate differs from the background only inside a bounded region (the object's footprint)
    changed = (full[:-1] != bg_only[None]).any(-1).flatten(1).sum(1)
    assert int(changed.max()) < res * res // 8 and int(changed[-1]) > 0
    # un-moved object: T_WO_2 == T_WO_1 -> virtual camera == real camera (convert_virtual_pose is the identity map)
    from dream2real_b200.reconstruction.combined_rendering import convert_virtual_pose
    T1 = accio2ngp.converter(scene["fg_pose"][None].astype(np.float64))[0]
    assert np.allclose(convert_virtual_pose(T1, T1, rp[0]), rp[0], atol=1e-12)


@pytest.mark.parametrize("res,rot90", [(800, True), (200, True), (96, False)])
def test_delta_preprocessing_is_bit_identical(tmp_path, res, rot90):
    """d2r_clip_preprocess_delta (background resized once + per-candidate affected windows) == d2r_clip_preprocess on
    the full frames, bit for bit, including a candidate whose rectangle is empty and one that covers the whole frame."""
    import torch
    from dream2real_b200 import synth
    from dream2real_b200.clip import ClipVision, make_hf_clip
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    d = str(tmp_path)
    scene = synth.make_scene("shopping", d, log2_hashmap_size=14, seed=5)
    tm = synth.SyntheticTaskModel(scene, "g", None, torch.device("cuda"))
    poses = _grid_poses(scene, [6, 6, 1, 1, 1, 1], list(range(0, 36, 5)))
    far = scene["fg_pose"].copy()[None].astype(np.float64)
    far[0, :3, 3] += [0.0, 3.0, 0.0]                      # sideways out of the frame: empty rectangle
    behind = scene["fg_pose"].copy()[None].astype(np.float64)
    behind[0, :3, 3] = scene["opt_cam_poses"][0][:3, 3]   # object around the camera: conservative full-frame rectangle
    poses = np.concatenate([poses, far, behind])
    r = renderer(d, tm, resolution=res)
    frames = r.render(accio2ngp.converter(poses), accio2ngp.converter(scene["opt_cam_poses"][:1]), [0], tm.depths[:1], tm.movable_masks,
                      save=False, return_tensor=True)
    rects = r.last_rects.cpu().numpy()
    print("rects", rects.tolist())
    assert (rects[-2, 2] < rects[-2, 0]) and tuple(rects[-1]) == (0, 0, res - 1, res - 1)      # the empty and the full-frame rectangle
    hf = make_hf_clip("ViT-B/32", seed=3)
    cv = ClipVision(hf, max_batch=16)
    full, _ = cv.preprocess(frames, rot90=rot90)
    full = full.clone()
    delta, _ = cv.preprocess(frames, rot90=rot90, bg_u8=r.last_bg_u8, rects=r.last_rects)
    assert torch.equal(full, delta)
    # the recorded background frame is what a frame without any object pixel looks like
    empty = np.nonzero(rects[:, 2] < rects[:, 0])[0]
    for k in empty:
        assert torch.equal(frames[int(k)], r.last_bg_u8)
    # and every frame equals it outside its rectangle
    for k in range(frames.shape[0]):
        x0, y0, x1, y1 = [int(v) for v in rects[k]]
        m = torch.ones(res, res, dtype=torch.bool, device=frames.device)
        if x1 >= x0 and y1 >= y0:
            m[y0:y1 + 1, x0:x1 + 1] = False
        assert torch.equal(frames[k][m], r.last_bg_u8[m])


def test_multi_view_scoring_is_the_mean_over_views(tmp_path):
    """render_cam_pose_idx with L > 1 -> one score per pose = mean over its L per-view scores."""
    import torch
    from dream2real_b200 import clip_scoring, synth
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from test_e2e_gpu import _tiny_clip
    d = str(tmp_path)
    scene = synth.make_scene("shopping", d, log2_hashmap_size=12, seed=4)
    tm = synth.SyntheticTaskModel(scene, "goal", ["norm"], torch.device("cuda"))
    model = _tiny_clip(11)
    ids = torch.randint(3, 900, (2, 6))
    ids[:, -1] = 2
    kw = dict(sample_res=[3, 3, 1, 1, 1, 1], phys_check=synth.all_valid_phys_check, scene_type=3, smoothing=False, clip_model=model,
              text_inputs={"input_ids": ids}, save_renders=False)
    r = renderer(d, tm, resolution=64)
    r.fix_mask_index = True
    per_view = []
    for v in (0, 1):
        _, _, s = clip_scoring.optimise_pose_grid(r, tm.depths[v:v + 1], [v], tm, d, **kw)
        per_view.append(s)
    _, _, both = clip_scoring.optimise_pose_grid(r, tm.depths[:2], [0, 1], tm, d, **kw)
    assert both.shape == per_view[0].shape
    assert torch.allclose(both, (per_view[0] + per_view[1]) / 2, rtol=1e-5, atol=1e-6)
    _, _, mx = clip_scoring.optimise_pose_grid(r, tm.depths[:2], [0, 1], tm, d, multi_view="max", **kw)
    assert torch.allclose(mx, torch.maximum(per_view[0], per_view[1]), rtol=1e-5, atol=1e-6)


def _frame_hashes(d, res=800, n=43):
    """per-candidate sha1 of the composited u8 frames of n shopping candidates (2^19 tables)"""
    import hashlib

    import torch
    from dream2real_b200 import synth
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.utils import accio2ngp
    scene = synth.make_scene("shopping", d, log2_hashmap_size=19, seed=1234)
    tm = synth.SyntheticTaskModel(scene, "g", None, torch.device("cuda"))
    poses = _grid_poses(scene, [64, 64, 1, 1, 1, 1], list(range(0, 4096, 4096 // n))[:n])
    r = renderer(d, tm, resolution=res, max_candidates_per_launch=16)
    frames = r.render(accio2ngp.converter(poses), accio2ngp.converter(scene["opt_cam_poses"][:1]), [0], tm.depths[:1], tm.movable_masks,
                      save=False, return_tensor=True).cpu().numpy()
    return [hashlib.sha1(f.tobytes()).hexdigest() for f in frames]


def test_split_kernels_give_identical_frames(tmp_path):
    """The A/B partner of the default kernel (D2R_MARCH=split: k_gather_round / k_mlp_round, the round-1 path) runs the same
    arithmetic per sample: bit-identical frames at the bench configuration, and green golden-render parity on its own.
    The mode is read once per process, so the split side runs in a child process."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    mine = _frame_hashes(str(tmp_path / "a"))
    env = dict(os.environ, D2R_MARCH="split")
    out = str(tmp_path / "split.json")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "hashes", str(tmp_path / "b"), out], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    theirs = json.load(open(out))
    assert len(mine) == len(theirs) and len(set(mine)) > 1
    assert mine == theirs, [i for i, (a, b) in enumerate(zip(mine, theirs)) if a != b]
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_render_gpu.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


if __name__ == "__main__":
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if len(sys.argv) == 4 and sys.argv[1] == "hashes":
        json.dump(_frame_hashes(sys.argv[2]), open(sys.argv[3], "w"))

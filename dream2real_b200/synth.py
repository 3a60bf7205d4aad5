"""Seeded synthetic stand-ins for Dream2Real scenes (no network: the real `method_out/<scene>`
snapshots are HuggingFace downloads).  Each scene keeps the real scene's *geometry of the problem*
(SURVEY.md 8(d)): camera block and NeRF dataset transform of reference configs/*_demo.json:48-67,
`aabb_scale 2` (utils/accio2ngp.py:60), scene_centre, pose bounds / sample_res ordering of
vision_3d/obj_pose_opt.py; what is synthetic is the content: analytic occupancy (spheres, boxes,
a table slab) written into the density grid and random fp16 hash tables / MLPs whose density head
is calibrated so that a ray saturates within a few tens of samples like a trained surface.

Produces real `.ingp` files (fg_base.ingp / bg_base.ingp) that both this library and the reference's
pyngp load, plus minimal SceneModel / TaskModel stand-ins exposing the attributes the path reads
(scene_model.py:13-125).
"""
from __future__ import annotations

import os
import types
from typing import Dict, List, Optional

import numpy as np

from . import ingp

# reference configs/shopping_demo.json:48-67 (all demo scenes share this RealSense colour camera)
CAMERA = dict(fl_x=924.66912, fl_y=926.49735, k1=0.096692, k2=-0.166479, p1=-0.000194, p2=0.002049,
              cx=654.51953, cy=355.18523, w=1280, h=720)
DATASET_SCALE = 1.0
DATASET_OFFSET = (0.0, 0.3, 0.5)
AABB_SCALE = 2
SCENE_CENTRE = (0.5, 0.0, 0.035)

# NeRF-convention camera-to-world of the render view (tests/golden/kat_cam.json "nerf_cam")
NERF_CAM = np.array([[0.0, -0.8, 0.6, 0.95], [1.0, 0.0, 0.0, 0.02], [0.0, 0.6, 0.8, 0.55], [0.0, 0.0, 0.0, 1.0]])

SCENES = {
    # name: scene_type, fg radius, fg initial position (world), background primitives (world coordinates)
    "shopping": dict(scene_type=3, fg_r=0.04, fg_pos=(0.62, 0.08, 0.04),
                     boxes=[((0.30, -0.22, 0.0), (0.42, -0.10, 0.10)), ((0.55, -0.25, 0.0), (0.66, -0.16, 0.07)),
                            ((0.36, 0.05, 0.0), (0.47, 0.16, 0.06)), ((0.70, -0.05, 0.0), (0.78, 0.03, 0.12)),
                            ((0.25, -0.02, 0.0), (0.31, 0.04, 0.09)), ((0.52, 0.18, 0.0), (0.60, 0.26, 0.05))], spheres=[]),
    "pool_triangle": dict(scene_type=0, fg_r=0.028, fg_pos=(0.60, 0.10, 0.028), boxes=[],
                          spheres=[((0.40 + 0.05 * (i % 5), -0.08 + 0.05 * (i // 5), 0.028), 0.028) for i in range(15)]),
    "shelf": dict(scene_type=1, fg_r=0.035, fg_pos=(0.45, 0.30, 0.10), cam_shift=(0.0, 0.40, 0.15),
                  boxes=[((0.30, 0.36, 0.0), (0.75, 0.50, 0.03)), ((0.30, 0.36, 0.20), (0.75, 0.50, 0.23)),
                         ((0.30, 0.36, 0.40), (0.75, 0.50, 0.43)), ((0.30, 0.48, 0.0), (0.75, 0.50, 0.45))], spheres=[]),
    "synthetic8": dict(scene_type=1, fg_r=0.04, fg_pos=(0.60, 0.20, 0.06), cam_shift=(0.0, 0.40, 0.15),
                       boxes=[((0.30 + 0.11 * i, 0.36, 0.0), (0.38 + 0.11 * i, 0.44, 0.06 + 0.03 * i)) for i in range(4)],
                       spheres=[((0.32 + 0.12 * i, 0.30, 0.05), 0.04) for i in range(4)]),
}


def world_to_ngp(p):
    """NerfDataset::nerf_position_to_ngp (nerf_loader.h:148-151): p*scale + offset, then xyz <- yzx."""
    p = np.asarray(p, np.float64) * DATASET_SCALE + np.asarray(DATASET_OFFSET)
    return p[..., [1, 2, 0]]


def _cell_centres(cascade: int):
    """NGP-space centres of the 128^3 cells of a cascade, in Morton order."""
    idx = np.arange(128 ** 3, dtype=np.uint32)

    def inv(x):
        x = x & np.uint32(0x49249249)
        x = (x | (x >> np.uint32(2))) & np.uint32(0xc30c30c3)
        x = (x | (x >> np.uint32(4))) & np.uint32(0x0f00f00f)
        x = (x | (x >> np.uint32(8))) & np.uint32(0xff0000ff)
        x = (x | (x >> np.uint32(16))) & np.uint32(0x0000ffff)
        return x
    xyz = np.stack([inv(idx), inv(idx >> np.uint32(1)), inv(idx >> np.uint32(2))], -1).astype(np.float64)
    size = 2.0 ** cascade
    return 0.5 - size / 2 + size * (xyz + 0.5) / 128.0, size / 128.0


def _occupancy(spheres_ngp, boxes_ngp, n_cascades: int) -> np.ndarray:
    """Analytic density grid: +1 in every cell that touches a primitive, -1 elsewhere (fp16)."""
    out = np.full(n_cascades * 128 ** 3, -1.0, np.float16)
    for c in range(n_cascades):
        centres, cell = _cell_centres(c)
        occ = np.zeros(centres.shape[0], bool)
        half_diag = cell * 0.5 * np.sqrt(3.0)
        for (ctr, r) in spheres_ngp:
            occ |= np.linalg.norm(centres - np.asarray(ctr), axis=-1) <= r + half_diag
        for (lo, hi) in boxes_ngp:
            lo, hi = np.minimum(lo, hi) - cell * 0.5, np.maximum(lo, hi) + cell * 0.5
            occ |= np.all((centres >= lo) & (centres <= hi), axis=-1)
        out[c * 128 ** 3:(c + 1) * 128 ** 3][occ] = 1.0
    return out


def _hash_encode_f32(grid: ingp.GridConfig, table: np.ndarray, x01: np.ndarray) -> np.ndarray:
    """fp32 hash-grid lookup, only used to calibrate the density head of a synthetic model."""
    n = x01.shape[0]
    out = np.zeros((n, grid.n_levels * 4), np.float32)
    for lvl in range(grid.n_levels):
        size = int(grid.offsets[lvl + 1] - grid.offsets[lvl])
        res = int(grid.resolutions[lvl])
        t = table[int(grid.offsets[lvl]): int(grid.offsets[lvl + 1])].astype(np.float32)
        pos = x01 * grid.scales[lvl] + 0.5
        fl = np.floor(pos)
        w = pos - fl
        pg = fl.astype(np.int64)
        acc = np.zeros((n, 4), np.float32)
        for corner in range(8):
            o = np.array([(corner >> d) & 1 for d in range(3)])
            c = (pg + o).astype(np.uint64)
            wt = np.prod(np.where(o == 1, w, 1 - w), axis=1)
            stride, idx = 1, np.zeros(n, np.uint64)
            for d in range(3):
                if stride > size:
                    break
                idx = idx + c[:, d] * np.uint64(stride)
                stride *= res
            if size < stride:
                idx = ((c[:, 0] * np.uint64(1)) ^ (c[:, 1] * np.uint64(2654435761)) ^ (c[:, 2] * np.uint64(805459861))) & np.uint64(0xFFFFFFFF)
            idx = idx % np.uint64(size)
            acc += wt[:, None].astype(np.float32) * t[idx.astype(np.int64)]
        out[:, lvl * 4:(lvl + 1) * 4] = acc
    return out


def _random_model(rng: np.random.Generator, log2_hashmap_size: int, sample_pts01: np.ndarray, target_density: float):
    pls = ingp.per_level_scale_for(AABB_SCALE, 16, 8)
    grid = ingp.grid_geometry(8, 4, log2_hashmap_size, 16, float(pls))
    table = rng.uniform(-0.5, 0.5, size=(int(grid.offsets[-1]), 4)).astype(np.float16)

    def xavier(o, i):
        return (rng.standard_normal((o, i)) * np.sqrt(2.0 / (i + o))).astype(np.float16)
    d0, d1 = xavier(64, 32), xavier(16, 64)
    c0, c1, c2 = xavier(64, 32), xavier(64, 64), xavier(16, 64)
    c2 = (c2.astype(np.float32) * 4.0).astype(np.float16)          # spread colours over the logistic
    # calibrate the raw density (row 0 of the density MLP output) to ~target on occupied points
    enc = _hash_encode_f32(grid, table, sample_pts01.astype(np.float32))
    h = np.maximum(enc @ d0.astype(np.float32).T, 0)
    d1f = d1.astype(np.float32)
    d1f[0] = np.abs(d1f[0])                                          # positive head on ReLU features -> density > 0
    raw = h @ d1f[0]
    d1f[0] *= target_density / max(float(np.median(raw)), 1e-6)
    d1 = d1f.astype(np.float16)
    params = np.concatenate([d0.ravel(), d1.ravel(), c0.ravel(), c1.ravel(), c2.ravel(), table.ravel()]).astype(np.float16)
    return params, grid


def make_scene(name: str, out_dir: str, log2_hashmap_size: int = 19, seed: int = 1234, n_views: int = 2,
               target_density: float = 4.0) -> Dict:
    """Write {out_dir}/fg_base.ingp and bg_base.ingp and return the host-side scene description."""
    spec = SCENES[name]
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    n_casc = 2  # aabb_scale 2 -> cascades 0 and 1
    views = [dict(CAMERA) for _ in range(n_views)]
    cams_nerf = []
    for v in range(n_views):
        m = NERF_CAM.copy()
        m[:3, 3] += np.asarray(spec.get("cam_shift", (0.0, 0.0, 0.0)))   # shelf scenes: the camera faces the shelf (y ~ 0.42)
        m[1, 3] += 0.06 * v          # second view: small sideways baseline
        cams_nerf.append(m)
    fg_pos = np.asarray(spec["fg_pos"], np.float64)
    # foreground: the movable object at its initial pose
    fg_grid = _occupancy([(world_to_ngp(fg_pos), spec["fg_r"] * DATASET_SCALE)], [], n_casc)
    pts = world_to_ngp(fg_pos + rng.uniform(-1, 1, (512, 3)) * spec["fg_r"] * 0.5)
    fg_params, _ = _random_model(rng, log2_hashmap_size, (pts + 0.5) / 2.0, target_density)
    cfg = ingp.build_snapshot_config(fg_params, fg_grid, aabb_scale=AABB_SCALE, log2_hashmap_size=log2_hashmap_size,
                                     views=views, xforms=cams_nerf, scale=DATASET_SCALE, offset=DATASET_OFFSET,
                                     background_color=(0.0, 0.0, 0.0, 0.0))   # train_ngp.py:72
    ingp.save_snapshot(os.path.join(out_dir, "fg_base.ingp"), cfg)
    # background: table slab + primitives (the movable object is masked out of the bg model)
    boxes = [(world_to_ngp(np.array(lo)), world_to_ngp(np.array(hi))) for lo, hi in spec["boxes"]]
    boxes.append((world_to_ngp(np.array((0.15, -0.45, -0.03))), world_to_ngp(np.array((0.95, 0.60, 0.0)))))   # table
    spheres = [(world_to_ngp(np.array(c)), r * DATASET_SCALE) for c, r in spec["spheres"]]
    bg_grid = _occupancy(spheres, boxes, n_casc)
    pts = world_to_ngp(np.stack([rng.uniform(0.2, 0.9, 512), rng.uniform(-0.4, 0.5, 512), rng.uniform(-0.02, 0.0, 512)], -1))
    bg_params, _ = _random_model(rng, log2_hashmap_size, (pts + 0.5) / 2.0, target_density)
    cfg = ingp.build_snapshot_config(bg_params, bg_grid, aabb_scale=AABB_SCALE, log2_hashmap_size=log2_hashmap_size,
                                     views=views, xforms=cams_nerf, scale=DATASET_SCALE, offset=DATASET_OFFSET,
                                     background_color=(0.0, 0.0, 0.0, 0.0))
    ingp.save_snapshot(os.path.join(out_dir, "bg_base.ingp"), cfg)

    # OpenCV-convention camera poses (what scene_model.opt_cam_poses holds; converter() flips y,z back)
    cams_cv = []
    for m in cams_nerf:
        c = m.copy()
        c[:3, 1] *= -1
        c[:3, 2] *= -1
        cams_cv.append(c)
    # sensor depth: z-depth of the table plane (world z = 0) through the pinhole model, metres
    H, W = CAMERA["h"], CAMERA["w"]
    ys, xs = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
    depths, masks = [], []
    for c in cams_cv:
        d_cam = np.stack([(xs - CAMERA["cx"]) / CAMERA["fl_x"], (ys - CAMERA["cy"]) / CAMERA["fl_y"], np.ones_like(xs)], -1)
        d_w = d_cam @ c[:3, :3].T
        with np.errstate(divide="ignore", invalid="ignore"):
            t = -c[2, 3] / d_w[..., 2]
        z = np.where((t > 0) & np.isfinite(t), t, 0.0)      # ray parameter == z-depth because d_cam.z == 1
        depths.append(np.clip(z, 0, 4.0).astype(np.float16))
        # True = NOT the movable object (scene_model.py:47): project the fg sphere
        pc = (fg_pos - c[:3, 3]) @ c[:3, :3]
        u, v = CAMERA["fl_x"] * pc[0] / pc[2] + CAMERA["cx"], CAMERA["fl_y"] * pc[1] / pc[2] + CAMERA["cy"]
        rad = CAMERA["fl_x"] * spec["fg_r"] / pc[2] * 1.2
        masks.append(~((xs - u) ** 2 + (ys - v) ** 2 <= rad ** 2))
    return dict(name=name, dir=out_dir, scene_type=spec["scene_type"], scene_centre=np.array(SCENE_CENTRE, np.float32),
                fg_pose=np.block([[np.eye(3), fg_pos[:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]),
                opt_cam_poses=np.stack(cams_cv), depths=np.stack(depths), movable_masks=np.stack(masks),
                log2_hashmap_size=log2_hashmap_size)


class _Obj:
    """ObjectModel stand-in (scene_model.py:13-24): .vis_model, .pose (torch [4,4]), .name"""

    def __init__(self, vis_model, pose, name):
        self.vis_model, self.pose, self.name = vis_model, pose, name


class SyntheticTaskModel:
    """TaskModel stand-in exposing what the path reads (scene_model.py:40-125)."""

    def __init__(self, scene: Dict, goal_caption: str, norm_captions: Optional[List[str]], device):
        import torch
        from .reconstruction.ngp_visual_model import get_vis_ngps
        self.scene = scene
        self.goal_caption, self.norm_captions = goal_caption, norm_captions
        self.user_instr = goal_caption
        self.scene_model = types.SimpleNamespace(
            scene_centre=torch.tensor(scene["scene_centre"]), device=device, scene_type=scene["scene_type"],
            opt_cam_poses=[torch.tensor(p) for p in scene["opt_cam_poses"]])
        self.movable_masks = torch.from_numpy(scene["movable_masks"]).to(device)
        self.depths = torch.from_numpy(scene["depths"]).to(device)
        fg = get_vis_ngps(None, None, scene["scene_type"], use_cache=True, data_dir=scene["dir"], fg=True)
        bg = get_vis_ngps(None, None, scene["scene_type"], use_cache=True, data_dir=scene["dir"], fg=False)
        self.movable_obj = _Obj(fg, torch.tensor(scene["fg_pose"], dtype=torch.float32), "movable")
        self.task_bground_obj = _Obj(bg, torch.eye(4), "task_background")
        self.topdown = False

    def free_visual_models(self):
        """scene_model.py free_visual_models: the reference frees the NeRFs to make room for CLIP; with
        180 GB of HBM3e both stay resident (they are ~25 MB each) so repeated queries skip the reload."""
        return None


def all_valid_phys_check(pose_batch, task_model, valid_so_far):
    """phys_check stand-in = dream2real.py:325 when physics checks are disabled."""
    return valid_so_far

"""CPU oracle of the pre-render physics filter: which candidate poses are collision-free, supported and stable.

TEST INFRASTRUCTURE ONLY (see oracle/ngp_oracle.py header): only tests/ and bench.py's checker legs import it.

What it restates: the control flow of `unsupcol_check` in reference vision_3d/physics_utils.py:248-378 --
    orientation-uniqueness mask over the rotations of the first position (:260-281), the regrasp mask for the embodied
    setting (:284-303), and per remaining pose: in collision -> invalid (:317-326); moved 2 cm down along gravity it must touch
    something unless the pose lies below the table plane (:329-343); and, lowered, it must still touch something when pushed
    4 cm along +-x and +-y (:351-368).
What it cannot restate: the collision primitive.  The reference asks pybullet (`pyb_planner.pairwise_collision` on
    GEOM_MESH collision shapes built from Poisson-reconstructed meshes, :232-246); pybullet and the meshes are not available
    here (requirements.txt: pybullet==3.2.5, un-vendored), so PARITY WITH PYBULLET IS UNPINNED.  The primitive used instead is
    the one the GPU kernel implements, stated here independently in numpy:
        collide(pose, shift) <=> some occupied cell centre of the movable object's NeRF density grid, moved by
                                 pose . init_pose^-1 and then by `shift`, lands in an occupied cell of the background NeRF's
                                 density grid (the same multi-cascade bitfield lookup the renderer's DDA uses,
                                 nerf_device.cuh:430-447 / ngp_oracle.density_grid_occupied_at).
"""
from __future__ import annotations

import numpy as np

from . import ngp_oracle as O

GRAVITY_DIRECTION = np.array([0.0, 0.0, -1.0])      # physics_utils.py:18


def occupied_points_ngp(bitfield: np.ndarray, max_cascade: int) -> np.ndarray:
    """Centres (NGP coordinates) of the occupied cells: cascade 0 everywhere, cascade c > 0 only outside cascade c-1's cube."""
    n = O.NERF_GRID_N_CELLS
    pts = []
    for c in range(max_cascade + 1):
        cells = np.nonzero(np.unpackbits(bitfield[c * (n // 8):(c + 1) * (n // 8)], bitorder="little"))[0].astype(np.uint32)
        if cells.size == 0:
            continue
        xyz = np.stack([O.morton3D_invert(cells), O.morton3D_invert(cells >> np.uint32(1)), O.morton3D_invert(cells >> np.uint32(2))], 1)
        size = 2.0 ** c
        p = 0.5 - size / 2 + size * (xyz.astype(np.float64) + 0.5) / 128.0
        if c > 0:
            inner = 2.0 ** (c - 1)
            p = p[np.any(np.abs(p - 0.5) > inner / 2, axis=1)]
        pts.append(p)
    return np.concatenate(pts, 0).astype(np.float32) if pts else np.zeros((0, 3), np.float32)


def ngp_to_world(p_ngp, scale, offset):
    """inverse of NerfDataset::nerf_position_to_ngp (nerf_loader.h:148-151): xyz <- zxy, then (p - offset) / scale"""
    p = np.asarray(p_ngp, np.float32)[..., [2, 0, 1]]
    return ((p - np.asarray(offset, np.float32)) / np.float32(scale)).astype(np.float32)


def world_to_ngp(p_world, scale, offset):
    p = np.asarray(p_world, np.float32) * np.float32(scale) + np.asarray(offset, np.float32)
    return p[..., [1, 2, 0]].astype(np.float32)


def collide(points_world, rel_34, shift, bg_bitfield, bg_max_cascade, scale, offset):
    """the occupancy-overlap primitive for one pose: points_world [n,3] f32, rel_34 [3,4] f32 (pose . init^-1), shift [3]"""
    R, t = rel_34[:, :3].astype(np.float32), rel_34[:, 3].astype(np.float32)
    # same operation order as the kernel: p' = R p + t, row by row in fp32, then + shift
    p = (points_world[:, 0:1] * R[:, 0] + points_world[:, 1:2] * R[:, 1] + points_world[:, 2:3] * R[:, 2] + t).astype(np.float32)
    p = (p + np.asarray(shift, np.float32)).astype(np.float32)
    q = world_to_ngp(p, scale, offset)
    mip = np.minimum(O.mip_from_pos(q), bg_max_cascade).astype(np.uint32)
    return bool(np.any(O.density_grid_occupied_at(q, bg_bitfield, mip)))


def orientation_masks(pose_batch, sample_res, valid_so_far, disallow_regrasp):
    """physics_utils.py:260-303 -> (uniqueness mask [N], regrasp mask [N]) as bool arrays (restated with numpy)."""
    P = np.asarray(pose_batch, np.float32).reshape(-1, 4, 4)
    per_pos = int(sample_res[3] * sample_res[4] * sample_res[5])
    n_pos = int(sample_res[0] * sample_res[1] * sample_res[2])
    first = P[:per_pos, :3, :3]
    uniq = np.ones(per_pos, bool)
    seen = []
    for i in range(per_pos):
        if any(np.all(np.isclose(first[i], s, atol=0.01, rtol=1e-5)) for s in seen):      # torch.isclose(atol=0.01): rtol default 1e-5
            uniq[i] = False
        else:
            seen.append(first[i])
    uniq_all = np.tile(uniq, n_pos)
    v = np.asarray(valid_so_far, bool) & uniq_all
    regrasp = np.ones(per_pos, bool)
    if disallow_regrasp:
        for i in range(per_pos):
            if not v[i]:
                regrasp[i] = False
                continue
            z = first[i][:, 2]
            if not (z @ np.array([0, 0, 1.0], np.float32) > 0.9 or z @ np.array([0, -1.0, 0], np.float32) > 0.9):
                regrasp[i] = False
    return uniq_all, np.tile(regrasp, n_pos)


def unsupcol_check(pose_batch, init_pose, points_world, bg_bitfield, bg_max_cascade, scale, offset, scene_centre_z, sample_res,
                   valid_so_far, disallow_regrasp=False, unsup_thresh=0.02, stability_check=True, p_dist=0.04):
    """physics_utils.py:248-378 with the occupancy-overlap primitive.  Returns bool [N]."""
    P = np.asarray(pose_batch, np.float32).reshape(-1, 4, 4)
    uniq, regrasp = orientation_masks(P, sample_res, valid_so_far, disallow_regrasp)
    valid = np.asarray(valid_so_far, bool) & uniq & regrasp
    rel = (P.astype(np.float64) @ np.linalg.inv(np.asarray(init_pose, np.float64))).astype(np.float32)[:, :3, :]
    lower = (unsup_thresh * GRAVITY_DIRECTION).astype(np.float32)
    perturb = [np.array(v, np.float32) * np.float32(p_dist) for v in ([1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0])]
    args = (bg_bitfield, bg_max_cascade, scale, offset)
    for i in range(P.shape[0]):
        if not valid[i]:
            continue
        if collide(points_world, rel[i], np.zeros(3, np.float32), *args):
            valid[i] = False
            continue
        below_table = P[i, 2, 3] < np.float32(scene_centre_z)
        if below_table:
            continue
        if not collide(points_world, rel[i], lower, *args):
            valid[i] = False
            continue
        if stability_check:
            for v in perturb:
                if not collide(points_world, rel[i], (lower + v).astype(np.float32), *args):
                    valid[i] = False
                    break
    return valid

"""Pre-render physics filter -- the B200 replacement for `create_unsupcol_check` of reference
vision_3d/physics_utils.py:232-378.

Same call surface: `create_unsupcol_check(pyb_planner, task_model, sample_res, embodied, ...)` returns
`(unsupcol_check, static_obj_handles, movable_handles)` and `unsupcol_check(pose_batch, task_model, valid_so_far)` returns the
bool mask `optimise_pose_grid` consumes as `phys_check` (clip_scoring.py:109-112).  The orientation-uniqueness and regrasp
masks are the reference's (they look at the handful of rotations of the first position only); the per-pose collision / support /
stability loop -- one pybullet round trip per query in the reference, 2.2 M poses for the shelf demo -- is ONE kernel launch
over the whole grid (csrc/d2r_phys.cu).

The collision primitive differs from the reference's: pybullet collides Poisson-reconstructed meshes (physics_utils.py:232-246),
which do not exist on this path; here the two NeRFs the path already holds are overlapped -- occupied cell centres of the movable
object's density grid against the background model's occupancy bitfield.  `pyb_planner` is accepted and ignored.
"""
import ctypes as C

import numpy as np
import torch

from .. import _native as N

GRAVITY_DIRECTION = np.array([0, 0, -1])      # physics_utils.py:18


def occupied_points_world(testbed):
    """Occupied cell centres of a loaded model's density grid in WORLD coordinates [n,3] float32: cascade 0 everywhere, cascade
    c > 0 only outside cascade c-1's cube (inverse of NerfDataset::nerf_position_to_ngp, nerf_loader.h:148-151)."""
    bits = testbed.occupancy_bitfield()
    snap = testbed.snapshot
    n = 128 ** 3

    def inv(x):      # morton3D_invert (tiny-cuda-nn common_device.h:773-785)
        x = x & np.uint32(0x49249249)
        x = (x | (x >> np.uint32(2))) & np.uint32(0xc30c30c3)
        x = (x | (x >> np.uint32(4))) & np.uint32(0x0f00f00f)
        x = (x | (x >> np.uint32(8))) & np.uint32(0xff0000ff)
        x = (x | (x >> np.uint32(16))) & np.uint32(0x0000ffff)
        return x
    pts = []
    for c in range(snap.max_cascade + 1):
        cells = np.nonzero(np.unpackbits(bits[c * (n // 8):(c + 1) * (n // 8)], bitorder="little"))[0].astype(np.uint32)
        if cells.size == 0:
            continue
        xyz = np.stack([inv(cells), inv(cells >> np.uint32(1)), inv(cells >> np.uint32(2))], 1)
        size = 2.0 ** c
        p = 0.5 - size / 2 + size * (xyz.astype(np.float64) + 0.5) / 128.0
        if c > 0:
            p = p[np.any(np.abs(p - 0.5) > 2.0 ** (c - 1) / 2, axis=1)]
        pts.append(p)
    p = np.concatenate(pts, 0).astype(np.float32) if pts else np.zeros((0, 3), np.float32)
    p = p[:, [2, 0, 1]]
    return ((p - np.asarray(snap.dataset_offset, np.float32)) / np.float32(snap.dataset_scale)).astype(np.float32)


def create_unsupcol_check(pyb_planner, task_model, sample_res, embodied, unsup_thresh=0.02, lazy_phys_mods=True, stability_check=True):
    """physics_utils.py:232-378.  Returns (unsupcol_check, static_obj_handles, movable_handles); the handle lists are empty (there
    is no pybullet world behind this check)."""
    fg, bg = task_model.movable_obj.vis_model, task_model.task_bground_obj.vis_model
    dev = torch.device("cuda", bg.device)
    pts = occupied_points_world(fg)
    pts4 = torch.from_numpy(np.concatenate([pts, np.zeros((pts.shape[0], 1), np.float32)], 1)).to(dev).contiguous()
    cfg = N.PhysCfg()
    cfg.dataset_scale = float(bg.snapshot.dataset_scale)
    cfg.dataset_offset[:] = [float(v) for v in bg.snapshot.dataset_offset]
    cfg.scene_centre_z = float(task_model.scene_model.scene_centre[2])
    cfg.unsup_thresh, cfg.p_dist, cfg.stability_check = float(unsup_thresh), 0.04, 1 if stability_check else 0

    def unsupcol_check(pose_batch, task_model, valid_so_far, disallow_regrasp=embodied):
        valid_so_far = valid_so_far.clone()
        pose_batch = pose_batch.view(-1, 4, 4)
        n_pos = sample_res[0] * sample_res[1] * sample_res[2]
        # duplicate orientations with the same rotation matrix: decided on the first position, repeated for all (:260-281)
        sampled_oris_per_pos = sample_res[3] * sample_res[4] * sample_res[5]
        first_pos_oris = pose_batch[:sampled_oris_per_pos, :3, :3]
        first_pos_validity_mask = torch.ones(sampled_oris_per_pos, dtype=torch.bool, device=first_pos_oris.device)
        oris_seen_so_far = []
        for i in range(first_pos_oris.shape[0]):
            ori = first_pos_oris[i]
            if any(bool(torch.all(torch.isclose(ori, seen_ori, atol=0.01))) for seen_ori in oris_seen_so_far):
                first_pos_validity_mask[i] = 0
            else:
                oris_seen_so_far.append(ori)
        first_pos_validity_mask = first_pos_validity_mask.repeat(n_pos)
        valid_so_far &= first_pos_validity_mask.to(valid_so_far.device)
        print(f'Of {pose_batch.shape[0]} sampled poses, {first_pos_validity_mask.sum()} pass orientation uniqueness check.')
        # embodied: only orientations whose z axis faces the camera may be regrasped (:284-303)
        first_pos_validity_mask = torch.ones(sampled_oris_per_pos, dtype=torch.bool, device=first_pos_oris.device)
        if disallow_regrasp:
            for i in range(first_pos_oris.shape[0]):
                if valid_so_far[i] == 0:
                    first_pos_validity_mask[i] = 0
                    continue
                obj_z_vector = first_pos_oris[i][:, 2]
                ori_facing_cam = bool(obj_z_vector @ torch.tensor([0, 0, 1.0], device=obj_z_vector.device) > 0.9) or \
                    bool(obj_z_vector @ torch.tensor([0, -1.0, 0], device=obj_z_vector.device) > 0.9)
                if not ori_facing_cam:
                    first_pos_validity_mask[i] = 0
        first_pos_validity_mask = first_pos_validity_mask.repeat(n_pos)
        valid_so_far &= first_pos_validity_mask.to(valid_so_far.device)
        print(f'Of {pose_batch.shape[0]} sampled poses, {first_pos_validity_mask.sum()} also pass regrasp check.')

        print("Checking each pose for colliding, unsupported or unstable objects...")
        # transforms = pose . init_pose^-1 (:252-253), float64 on the host like the oracle, then one launch for all poses
        init = task_model.movable_obj.pose.detach().cpu().numpy().astype(np.float64)
        P64 = pose_batch.detach().cpu().numpy().astype(np.float32).astype(np.float64)
        rel = np.ascontiguousarray((P64 @ np.linalg.inv(init)).astype(np.float32)[:, :3, :].reshape(-1, 12))
        n = rel.shape[0]
        rel_d = torch.from_numpy(rel).to(dev)
        z_d = pose_batch[:, 2, 3].detach().to(device=dev, dtype=torch.float32).contiguous()
        vin = valid_so_far.to(device=dev, dtype=torch.uint8).contiguous()
        vout = torch.empty(n, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib().d2r_phys_check(bg._model, pts4.data_ptr(), int(pts4.shape[0]), rel_d.data_ptr(), z_d.data_ptr(), vin.data_ptr(), n,
                                           C.byref(cfg), vout.data_ptr(), N.stream_ptr()), "phys_check")
        return vout.bool().to(valid_so_far.device)

    return unsupcol_check, [], []

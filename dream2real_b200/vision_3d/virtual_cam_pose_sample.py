"""reference vision_3d/virtual_cam_pose_sample.py:4-8."""
import numpy as np


def get_virtual_cam_poses(task_model, render_cam_pose_idx):
    poses = task_model.scene_model.opt_cam_poses
    return np.stack([np.asarray(poses[idx].cpu() if hasattr(poses[idx], "cpu") else poses[idx]) for idx in render_cam_pose_idx], axis=0)

"""CPU oracle for the Python half of the path: pose grid, frame conversions, depth-test composite,
CLIP preprocessing, CLIP forward, score normalisation, heat-map smoothing.

TEST INFRASTRUCTURE ONLY (see oracle/ngp_oracle.py header).  Paths are relative to /root/reference.

Third-party pieces the reference calls and that are NOT vendored under /root/reference:
  * transformers==4.27.3 CLIPProcessor / CLIPModel (requirements.txt:256)  -> here: the installed
    transformers CLIPModel (eager attention, fp32) is the arbiter of the forward pass; the image
    processor is restated from 4.27.3 image_transforms (PIL BICUBIC resize on uint8 -> float64
    rescale 1/255 -> float32 -> (x-mean)/std) with the real Pillow doing the resampling.
  * pytorch3d euler_angles_to_matrix (vision_3d/obj_pose_opt.py:3) -> restated (Rx @ Ry @ Rz).
  * torchvision gaussian_blur / cv2.resize -> the real libraries are called.
Pinning: the reference ships no tests for these functions ("parity unpinned" by reference tests);
pins used instead: Pillow itself (resize), HF CLIPModel itself (forward), torchvision (blur), and
hand-checkable identities in tests/test_post_oracle.py.
"""
from __future__ import annotations

import numpy as np

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

# scene bounds of vision_3d/obj_pose_opt.py:16-36, keyed by scene_type
POSE_BOUNDS = {
    0: dict(x=(-0.12, 0.04), y=(-0.10, 0.06), z=(0.00, 0.085), rot=(0.0, 0.0)),            # pool table
    1: dict(x=(-0.15, 0.20), y=(0.40, 0.44), z=(0.04, 0.41), rot=(-np.pi, np.pi / 2)),     # shelf
    3: dict(x=(-0.19, 0.15), y=(-0.25, 0.10), z=(0.00, 0.14), rot=(0.0, 0.0)),             # shopping
}


def euler_xyz_to_matrix(e):
    """pytorch3d.transforms.euler_angles_to_matrix(e, 'XYZ') = Rx(e0) @ Ry(e1) @ Rz(e2)."""
    import torch
    cx, sx = torch.cos(e[:, 0]), torch.sin(e[:, 0])
    cy, sy = torch.cos(e[:, 1]), torch.sin(e[:, 1])
    cz, sz = torch.cos(e[:, 2]), torch.sin(e[:, 2])
    o, z = torch.ones_like(cx), torch.zeros_like(cx)
    Rx = torch.stack([o, z, z, z, cx, -sx, z, sx, cx], -1).reshape(-1, 3, 3)
    Ry = torch.stack([cy, z, sy, z, o, z, -sy, z, cy], -1).reshape(-1, 3, 3)
    Rz = torch.stack([cz, -sz, z, sz, cz, z, z, z, o], -1).reshape(-1, 3, 3)
    return Rx @ Ry @ Rz


def sample_poses_grid(scene_centre, sample_res, scene_type):
    """vision_3d/obj_pose_opt.py:8-54 -> float32 [N,16] (x slowest ... z-rotation fastest)."""
    import torch
    if scene_type not in POSE_BOUNDS:
        raise NotImplementedError("scene_type %d not implemented" % scene_type)
    b = POSE_BOUNDS[scene_type]
    c = torch.as_tensor(scene_centre, dtype=torch.float32)
    axes = []
    for i, key in enumerate("xyz"):
        lo, hi = torch.tensor(b[key], dtype=torch.float32) + c[i]
        axes.append(torch.linspace(lo, hi, sample_res[i]))
    for i in range(3):
        lo, hi = torch.tensor(b["rot"], dtype=torch.float32)
        axes.append(torch.linspace(lo, hi, sample_res[3 + i]))
    combos = torch.cartesian_prod(*axes)
    if combos.dim() == 1:
        combos = combos[None]
    poses = torch.eye(4).repeat(combos.shape[0], 1, 1)
    poses[:, :3, 3] = combos[:, :3]
    poses[:, :3, :3] = euler_xyz_to_matrix(combos[:, 3:])
    return poses.reshape(-1, 16)


def converter(T):
    """utils/accio2ngp.py:133-139: negate the y and z rotation columns (OpenCV -> NeRF camera axes)."""
    out = np.array(T, copy=True)
    out[..., :3, 1] *= -1
    out[..., :3, 2] *= -1
    return out


def convert_virtual_pose(T_WO_1, T_WO_2, T_WC_1):
    """reconstruction/combined_rendering.py:250-263: virtual camera that sees the object at its
    initial pose the way the real camera would see it at the candidate pose."""
    T_O2_O1 = np.linalg.inv(T_WO_2) @ T_WO_1
    T_O1_C1 = np.linalg.inv(T_WO_1) @ T_WC_1
    return T_WO_1 @ T_O2_O1 @ T_O1_C1


def linear_to_srgb_py(x):
    """NGP scripts/common.py:142-144."""
    limit = 0.0031308
    with np.errstate(invalid="ignore"):
        return np.where(x > limit, 1.055 * (x ** (1.0 / 2.4)) - 0.055, 12.92 * x)


def composite(bg_image, bg_depth0, fg_image, fg_depth0):
    """reconstruction/combined_rendering.py:133-155.  bg_image/fg_image float32 [H,W,4]; *_depth0
    the depth channel [H,W].  Returns uint8 [H,W,3]."""
    fg_d = np.array(fg_depth0, np.float32, copy=True)
    bg_d = np.array(bg_depth0, np.float32, copy=True)
    fg_d[fg_d < 0.05] = 100
    bg_d[bg_d < 0.05] = 100
    near = fg_d < bg_d
    cb = bg_image.copy()
    cb[near, :] = fg_image[near, :]
    img = np.copy(cb)
    img[..., 0:3] = np.divide(img[..., 0:3], img[..., 3:4], out=np.zeros_like(img[..., 0:3]), where=img[..., 3:4] != 0)
    img[..., 0:3] = linear_to_srgb_py(img[..., 0:3])
    img = (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    img[img[..., 3] < 130, :] = 0
    return img[..., :3]


def rectify(img, resolution, as_u8=False):
    """rectify_depth / rectify_mask (combined_rendering.py:166-209): centre crop to square, cv2 cubic resize."""
    import cv2
    img = np.asarray(img)
    h, w = img.shape[:2]
    img = img[(h - w) // 2:(h - w) // 2 + w, :] if h > w else img[:, (w - h) // 2:(w - h) // 2 + h]
    img = img.astype(np.uint8 if as_u8 else np.float32)
    return cv2.resize(img, (resolution[0], resolution[1]), interpolation=cv2.INTER_CUBIC)


def background_depth(depth_gt, movable_mask, resolution):
    """combined_rendering.py:107-110: sensor depth, with the pixels the movable object used to cover pushed to 100."""
    d = rectify(depth_gt, resolution)
    m = rectify(movable_mask, resolution, as_u8=True)
    d = d.copy()
    d[m == 0] = 100
    return d


def clip_preprocess(images_u8, size):
    """transformers 4.27.3 CLIPImageProcessor on uint8 HWC images (square): PIL BICUBIC resize to
    `size`, /255 (float64 -> float32), normalise; returns float32 [K,3,size,size]."""
    from PIL import Image
    mean = np.array(OPENAI_CLIP_MEAN, dtype=np.float32)
    std = np.array(OPENAI_CLIP_STD, dtype=np.float32)
    out = []
    for im in images_u8:
        assert im.shape[0] == im.shape[1], "square renders only"
        pil = Image.fromarray(im)
        if pil.size != (size, size):
            pil = pil.resize((size, size), resample=Image.BICUBIC)
        a = (np.asarray(pil) * (1 / 255)).astype(np.float32)
        a = (a - mean) / std
        out.append(a.transpose(2, 0, 1))
    return np.stack(out).astype(np.float32)


def clip_logits(hf_model, pixel_values, input_ids, attention_mask=None):
    """clip_scoring.py:180-181 -> logits_per_image [K, n_captions] (fp32, eager attention)."""
    import torch
    with torch.no_grad():
        out = hf_model(pixel_values=torch.as_tensor(pixel_values), input_ids=input_ids, attention_mask=attention_mask)
    return out.logits_per_image


def normalise_scores(all_logits, n_goal=1):
    """clip_scoring.py:187-203: goal / mean(normalising) (template means when n_goal > 1)."""
    if all_logits.shape[1] == n_goal:
        return all_logits.mean(dim=1)
    return all_logits[:, :n_goal].mean(dim=1) / all_logits[:, n_goal:].mean(dim=1)


def spatially_smooth_heatmap(pose_scores, sample_res, sigma=0.7):
    """vision_3d/geometry_utils.py:252-269 with the real torchvision ops."""
    import torch
    import torchvision.transforms.functional as TF
    s = pose_scores.clone()
    min_nonzero = torch.min(s[s != 0]).item()
    zero_idxs = torch.nonzero(s == 0, as_tuple=True)
    s[zero_idxs] = min_nonzero
    rest = sample_res[2] * sample_res[3] * sample_res[4] * sample_res[5]
    s = s.view(sample_res[0] * sample_res[1], rest).swapaxes(0, 1).unsqueeze(1)
    s = s.reshape(rest, 1, sample_res[0], sample_res[1])
    s = TF.pad(s, padding=1, fill=min_nonzero, padding_mode="constant")
    s = TF.gaussian_blur(s, kernel_size=3, sigma=sigma)[:, :, 1:-1, 1:-1]
    s = s.reshape(rest, 1, sample_res[0] * sample_res[1]).squeeze(1).swapaxes(0, 1).reshape(-1)
    s[zero_idxs] = 0
    return s.contiguous()

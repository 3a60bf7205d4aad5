"""Combined (movable object over task background) rendering -- the B200 replacement for reference
reconstruction/combined_rendering.py:54-263.

Same class / method names and argument meaning.  What changed underneath: the background is rendered
once per view, every candidate pose is one entry of a batched launch of the fused march kernel
(Shade + Depth in one pass) whose epilogue performs the depth-test composite, un-premultiply, sRGB,
u8 quantisation and alpha threshold in registers (combined_rendering.py:133-155) -- no per-candidate
device->host copies, no NumPy compositing.
"""
import os
import shutil

import numpy as np

from .. import pyngp as ngp  # noqa: F401
from ..utils import accio2ngp


class obj_nerf:
    """Unit-test helper of the reference (combined_rendering.py:36-51): a snapshot wrapped like an ObjectModel."""

    def __init__(self, nerf_file, pose=None):
        import torch
        self.vis_model = ngp.Testbed(ngp.TestbedMode.Nerf)
        self.vis_model.load_file(nerf_file)
        self.vis_model.snap_to_pixel_centers = True
        self.vis_model.nerf.render_min_transmittance = 1e-4
        self.vis_model.shall_train = False
        self.testbed = self.vis_model
        self.pose = torch.eye(4) if pose is None else pose


def convert_virtual_pose(T_WO_1, T_WO_2, T_WC_1):
    """Virtual camera T_WC_2 such that the object at its initial pose T_WO_1 is seen the way the real
    camera T_WC_1 would see it at the candidate pose T_WO_2: T_WO_1 . T_WO_2^-1 . T_WC_1
    (combined_rendering.py:250-263; batched over leading dims of T_WO_2)."""
    T_O2_O1 = np.linalg.inv(T_WO_2) @ T_WO_1
    T_O1_C1 = np.linalg.inv(T_WO_1) @ T_WC_1
    return T_WO_1 @ T_O2_O1 @ T_O1_C1


class renderer:
    def __init__(self, data_dir, task_model=None, resolution=336, max_candidates_per_launch=2048):
        self.root = data_dir
        self.resolution = [int(resolution), int(resolution)]     # combined_rendering.py:86 hard-codes 336
        self.max_candidates_per_launch = int(max_candidates_per_launch)
        self.fix_mask_index = False   # reference indexes movable_masks by loop index, not camera index (:109)
        if task_model is not None:
            self.bg_obj = task_model.task_bground_obj
            self.fg_obj = task_model.movable_obj
        else:
            self.bg_obj = obj_nerf(os.path.join(self.root, "bg_base.ingp"))
            self.fg_obj = obj_nerf(os.path.join(self.root, "fg_base.ingp"))
        self.out_render_path = os.path.join(self.root, "cb_render")
        os.makedirs(self.out_render_path, exist_ok=True)
        self.last_n_samples = 0
        self.last_rects = None        # int32 CUDA [K,4] / uint8 CUDA [H,W,3] of the last single-view render(return_tensor=True):
        self.last_bg_u8 = None        # outside its rectangle a frame equals the composited background (preprocessing exploits it)
        self.count_samples = False    # True: read the network-sample count back after every launch (one host sync each)

    # ---- background -----------------------------------------------------------------------------
    def render_background(self, cam_matrix, view_idx, depth_gt=None, movable_mask=None):
        """bg Shade render + bg depth for one view (combined_rendering.py:95-113) -> CUDA tensors."""
        import torch
        W, H = self.resolution
        bg = self.bg_obj.vis_model
        bg.set_camera_to_training_view(view_idx)
        bg.background_color = [0.0, 0.0, 0.0, 1.0]
        bg.render_ground_truth = False
        cam = np.asarray(cam_matrix, dtype=np.float64)[None, :3, :]
        want_depth = depth_gt is None
        shade, depth = bg.render_batch(cam, W, H, want_shade=True, want_depth=want_depth)
        bg_image = shade[0].contiguous()
        if want_depth:
            bg_depth = depth[0, :, :, 0].contiguous()
        else:
            # combined_rendering.py:104-111 uses channel 0 of the 4-channel rectified depth only: skip the repeat
            d = self._rectify_depth_1ch(depth_gt, self.resolution)
            m = self.rectify_mask(movable_mask, self.resolution)
            np.putmask(d, m == 0, np.float32(100))
            bg_depth = torch.from_numpy(np.ascontiguousarray(d)).to(bg_image.device)
        return bg_image, bg_depth

    # ---- the hot loop ---------------------------------------------------------------------------
    def iter_render(self, valid_poses, render_poses, render_cam_pose_idx, depths_gt=None, movable_masks=None, save=False,
                    chunk=None, out=None):
        """The streaming form of render(): yields (render_idx, lo, hi, frames, rects, bg_u8) per launch of at most `chunk`
        candidates -- frames uint8 CUDA [hi-lo,H,W,3] (a reused buffer unless `out` [K*L,H,W,3] is given: consume it, on the
        current stream, before asking for the next chunk), rects int32 CUDA [hi-lo,4] and bg_u8 uint8 CUDA [H,W,3]: outside
        its rectangle a frame equals bg_u8 (ClipVision.preprocess exploits it).  Only O(chunk) frames exist at any time, so
        the reference's full pose grids (70 000 poses for the shopping demo, configs/shopping_demo.json:28) stream through
        a few GB.  save=True streams the PNGs of combined_rendering.py:157-159 chunk by chunk."""
        import torch
        T_WO_1 = accio2ngp.converter(np.expand_dims(self.fg_obj.pose.cpu().numpy(), axis=0)).astype(np.float64)[0]
        valid_poses = np.asarray(valid_poses, dtype=np.float64).reshape(-1, 4, 4)
        K = len(valid_poses)
        W, H = self.resolution
        chunk = int(chunk) if chunk else self.max_candidates_per_launch
        save = save and len(render_cam_pose_idx) == 1
        if save:
            if os.path.exists(self.out_render_path):
                shutil.rmtree(self.out_render_path)
            os.makedirs(self.out_render_path)
        self.last_n_samples = 0
        fg = self.fg_obj.vis_model
        inv_T_WO_2 = None                                                          # [K,4,4], computed while the GPU renders the background
        dev = torch.device("cuda", fg.device)
        buf = None if out is not None else torch.empty((min(chunk, K), H, W, 3), dtype=torch.uint8, device=dev)
        rect_buf = torch.empty((min(chunk, K), 4), dtype=torch.int32, device=dev) if out is None else None
        for render_idx in range(len(render_cam_pose_idx)):
            view_idx = render_cam_pose_idx[render_idx]
            T_WC_1 = np.asarray(render_poses[render_idx], dtype=np.float64)
            mask_idx = view_idx if self.fix_mask_index else render_idx
            bg_image, bg_depth = self.render_background(
                T_WC_1, view_idx, None if depths_gt is None else depths_gt[render_idx],
                None if depths_gt is None else movable_masks[mask_idx])
            fg.set_camera_to_training_view(view_idx)
            fg.render_ground_truth = False
            if inv_T_WO_2 is None:
                inv_T_WO_2 = np.linalg.inv(valid_poses)
            # T_WC_2 = T_WO_1 . (T_WO_2^-1 . T_WO_1) . (T_WO_1^-1 . T_WC_1), same association as the reference
            cams = T_WO_1 @ (inv_T_WO_2 @ T_WO_1) @ (np.linalg.inv(T_WO_1) @ T_WC_1)
            bg_u8 = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
            rects_all = torch.empty((K, 4), dtype=torch.int32, device=dev) if out is not None else None
            for s in range(0, K, chunk):
                e = min(s + chunk, K)
                frames = out[render_idx * K + s: render_idx * K + e] if out is not None else buf[: e - s]
                rects = rects_all[s:e] if out is not None else rect_buf[: e - s]
                fg.render_composite_batch(cams[s:e, :3, :], W, H, bg_image, bg_depth, out_u8=frames, count_samples=self.count_samples,
                                          rects_out=rects, bg_u8_out=bg_u8 if s == 0 else None)
                if self.count_samples:
                    self.last_n_samples += fg.last_n_samples
                if save:
                    import cv2
                    host = frames.cpu().numpy()
                    for i in range(host.shape[0]):
                        cv2.imwrite(os.path.join(self.out_render_path, f"cb_rgb_{s + i:04d}.png"), cv2.cvtColor(host[i], cv2.COLOR_RGB2BGR))
                yield render_idx, s, e, frames, rects, bg_u8
            self.last_rects, self.last_bg_u8 = (rects_all, bg_u8) if (out is not None and len(render_cam_pose_idx) == 1) else (None, None)

    def render(self, valid_poses, render_poses, render_cam_pose_idx, depths_gt=None, movable_masks=None, save=True,
               return_tensor=False):
        """valid_poses [K,4,4] (NeRF convention), render_poses [L,4,4], render_cam_pose_idx list[int],
        depths_gt torch [L,Hs,Ws], movable_masks torch bool [n_views,Hs,Ws].
        Returns the reference's list of K*L uint8 [H,W,3] arrays (combined_rendering.py:73-163), or (return_tensor=True) one
        uint8 CUDA tensor [K*L,H,W,3] that never leaves the device.  Both materialise every frame, like the reference;
        optimise_pose_grid streams through iter_render instead."""
        import torch
        K = int(np.asarray(valid_poses).reshape(-1, 4, 4).shape[0])
        W, H = self.resolution
        L = len(render_cam_pose_idx)
        if return_tensor:
            out = torch.empty((K * L, H, W, 3), dtype=torch.uint8, device=torch.device("cuda", self.fg_obj.vis_model.device))
            for _ in self.iter_render(valid_poses, render_poses, render_cam_pose_idx, depths_gt, movable_masks, save=save, out=out):
                pass
            return out
        host = np.empty((K * L, H, W, 3), np.uint8)
        for render_idx, s, e, frames, _, _ in self.iter_render(valid_poses, render_poses, render_cam_pose_idx, depths_gt, movable_masks, save=save):
            host[render_idx * K + s: render_idx * K + e] = frames.cpu().numpy()
        return [host[i] for i in range(host.shape[0])]

    # ---- sensor depth / mask rectification (combined_rendering.py:166-209) -------------------------
    @staticmethod
    def _center_crop(img):
        h, w = img.shape[:2]
        if h > w:
            return img[(h - w) // 2:(h - w) // 2 + w, :]
        return img[:, (w - h) // 2:(w - h) // 2 + h]

    def _rectify_depth_1ch(self, depth_ori, resolution):
        import cv2
        import torch
        if getattr(depth_ori, "is_cuda", False):
            # crop and widen fp16 -> fp32 on the GPU (exact) before the copy to the host: numpy's software half
            # conversion of the frame costs more than the resize itself
            img = np.ascontiguousarray(self._center_crop(depth_ori).to(torch.float32).cpu().numpy())
        else:
            img = self._center_crop(depth_ori.cpu().numpy()).astype(np.float32)
        return cv2.resize(img, (resolution[0], resolution[1]), interpolation=cv2.INTER_CUBIC)

    def rectify_depth(self, depth_ori, resolution):
        return np.repeat(np.expand_dims(self._rectify_depth_1ch(depth_ori, resolution), axis=2), 4, axis=2)

    def rectify_mask(self, mask_ori, resolution):
        import cv2
        img = self._center_crop(mask_ori.cpu().numpy()).astype(np.uint8)
        return cv2.resize(img, (resolution[0], resolution[1]), interpolation=cv2.INTER_CUBIC)

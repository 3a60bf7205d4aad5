"""`pyngp.Testbed`-shaped handle over libd2r_b200 -- the slice of the pybind11 API that
Dream2Real's hot path touches (reference reconstruction/instant-ngp/src/python_api.cu:262-566;
call sites reconstruction/combined_rendering.py:41-51,98-130 and ngp_visual_model.py:24-28).

Same names, same argument meaning, same error behaviour (RuntimeError with the reference's
message for bad snapshots).  Rendering itself is the fused sm_100a kernel behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from . import ingp


class TestbedMode(enum.Enum):      # python_api.cu:267-273
    Nerf = 0
    Sdf = 1
    Image = 2
    Volume = 3
    None_ = 4


class RenderMode(enum.Enum):       # python_api.cu:283-292 (only Shade and Depth are on the path)
    AO = 0
    Shade = 1
    Normals = 2
    Positions = 3
    Depth = 4
    Distortion = 5
    Cost = 6
    Slice = 7


Shade = RenderMode.Shade
Depth = RenderMode.Depth
Nerf = TestbedMode.Nerf


def nerf_matrix_to_ngp(m34: np.ndarray, scale: float, offset: Sequence[float], from_mitsuba: bool = False) -> np.ndarray:
    """NerfDataset::nerf_matrix_to_ngp (nerf_loader.h:101-121), float32, accepts [...,3,4]."""
    m = np.array(m34, dtype=np.float32, copy=True)[..., :3, :4]
    m[..., :, 1] *= np.float32(-1)
    m[..., :, 2] *= np.float32(-1)
    m[..., :, 3] = m[..., :, 3] * np.float32(scale) + np.asarray(offset, np.float32)
    if from_mitsuba:
        m[..., :, 0] *= np.float32(-1)
        m[..., :, 2] *= np.float32(-1)
        return m
    return m[..., [1, 2, 0], :]


class _NerfState:
    """testbed.nerf.* attributes the path reads/writes (python_api.cu:578-599)."""

    def __init__(self):
        self.render_min_transmittance = 0.01     # testbed.h:766
        self.render_with_lens_distortion = False
        self.cone_angle_constant = 1.0 / 256.0
        self.render_gbuffer_hard_edges = False


class Testbed:
    """Drop-in for the `pyngp.Testbed` calls on the Dream2Real path."""

    def __init__(self, mode: TestbedMode = TestbedMode.None_, device: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("dream2real_b200.Testbed needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.mode = mode
        self.background_color = [0.0, 0.0, 0.0, 1.0]       # testbed.h default m_background_color
        self.render_ground_truth = False
        self.render_groundtruth = False
        self.render_mode = RenderMode.Shade
        self.snap_to_pixel_centers = False
        self.shall_train = False
        self.root_dir = ""
        self.exposure = 0.0
        self.nerf = _NerfState()
        self.snapshot: Optional[ingp.Snapshot] = None
        self._model = C.c_void_p()
        self._views: Dict[Tuple[int, int, int, bool], C.c_void_p] = {}
        self._view_idx = 0
        self._lens_from_view = False
        self._camera_ngp = None      # [3,4] float32, NGP convention

    # ---- loading ---------------------------------------------------------------------------------
    def load_snapshot(self, path: str) -> None:
        """Testbed::load_snapshot (testbed.cu:4757-4880): RuntimeError on a bad file."""
        self._free_native()
        snap = ingp.load_snapshot(str(path))
        self.snapshot = snap
        self.mode = TestbedMode.Nerf
        self.background_color = [float(c) for c in snap.background_color]      # testbed.cu:4823
        self.exposure = snap.exposure
        if float(snap.exposure) != 0.0:
            # tonemap_kernel scales by 2^exposure (render_buffer.cu:529-561); Dream2Real snapshots are saved with exposure 0
            raise RuntimeError(f"snapshot exposure {snap.exposure} != 0 is not on the Dream2Real path")
        self.nerf.cone_angle_constant = snap.cone_angle_constant
        self._camera_ngp = snap.snapshot_camera.astype(np.float32)
        if snap.density_grid.size == 0:
            raise RuntimeError("snapshot has an empty density grid (untrained model): nothing to render")
        cfg = N.ModelCfg()
        g = snap.grid
        cfg.n_levels, cfg.n_features_per_level = g.n_levels, g.n_features_per_level
        cfg.log2_hashmap_size, cfg.base_resolution = g.log2_hashmap_size, g.base_resolution
        cfg.per_level_scale = g.per_level_scale
        cfg.max_cascade = snap.max_cascade
        cfg.aabb_min[:] = snap.aabb_min.tolist()
        cfg.aabb_max[:] = snap.aabb_max.tolist()
        cfg.render_aabb_min[:] = snap.render_aabb_min.tolist()
        cfg.render_aabb_max[:] = snap.render_aabb_max.tolist()
        cfg.render_aabb_to_local[:] = snap.render_aabb_to_local.reshape(-1).tolist()
        cfg.cone_angle_constant = snap.cone_angle_constant
        cfg.min_transmittance = self.nerf.render_min_transmittance
        cfg.depth_scale = 1.0 / snap.dataset_scale
        params = np.ascontiguousarray(snap.params)
        grid = np.ascontiguousarray(snap.density_grid, dtype=np.float32)
        self._cfg = cfg
        N.check(N.lib().d2r_model_load(params.ctypes.data, params.size, grid.ctypes.data, grid.size,
                                       C.byref(cfg), self.device, C.byref(self._model)), "load_snapshot")
        self._loaded_min_T = self.nerf.render_min_transmittance

    def load_file(self, path: str) -> None:
        """Testbed::load_file (testbed.cu:305-370): only snapshots are on this path."""
        p = str(path)
        if p.endswith(".ingp") or p.endswith(".msgpack"):
            self.load_snapshot(p)
        else:
            raise RuntimeError(f"File '{p}': only .ingp/.msgpack snapshots can be loaded on this path "
                               "(NeRF training data loading is out of scope)")

    # ---- camera ----------------------------------------------------------------------------------
    def set_camera_to_training_view(self, trainview: int) -> None:
        """testbed.cu:453-468: intrinsics + lens of the view; render_with_lens_distortion := True."""
        self._require()
        if not 0 <= int(trainview) < len(self.snapshot.views):
            raise RuntimeError(f"training view {trainview} does not exist")
        self._view_idx = int(trainview)
        self._lens_from_view = True
        self.nerf.render_with_lens_distortion = True

    def set_nerf_camera_matrix(self, cam) -> None:
        """testbed.cu:401-403; cam is a [3,4] NeRF-convention camera-to-world matrix."""
        self._require()
        m = np.asarray(cam, dtype=np.float64).reshape(3, 4)
        s = self.snapshot
        self._camera_ngp = nerf_matrix_to_ngp(m, s.dataset_scale, s.dataset_offset, s.from_mitsuba)

    # ---- rendering -------------------------------------------------------------------------------
    def render(self, width: int, height: int, spp: int = 1, linear: bool = True, *a, **kw) -> np.ndarray:
        """Testbed::render_to_cpu (python_api.cu:123-201) -> float32 [h,w,4] linear premultiplied."""
        if spp != 1 or not linear:
            raise RuntimeError("only spp=1, linear=True renders are on the Dream2Real path")
        if self.render_mode not in (RenderMode.Shade, RenderMode.Depth, RenderMode.Cost):
            raise RuntimeError(f"render_mode {self.render_mode} is not on the Dream2Real path")
        if self.render_mode == RenderMode.Cost:
            # shade_kernel_nerf (testbed_nerf.cu:1322-1326): rgb = n_steps / 128, alpha 1 on the rays it keeps; the rest of
            # the frame is the tonemap background blend
            rgba, _, cost = self.render_batch(self._camera_ngp[None], width, height, ngp_convention=True, want_shade=True,
                                              want_depth=False, want_cost=True)
            out, n = rgba[0].cpu().numpy(), cost[0].cpu().numpy()
            kept = n > 0
            out[kept] = np.stack([n[kept] / 128.0] * 3 + [np.ones_like(n[kept])], -1)
            return out
        rgba, depth = self.render_batch(self._camera_ngp[None], width, height, ngp_convention=True,
                                        want_shade=self.render_mode == RenderMode.Shade,
                                        want_depth=self.render_mode == RenderMode.Depth)
        out = rgba if self.render_mode == RenderMode.Shade else depth
        return out[0].cpu().numpy()

    def view_handle(self, width: int, height: int) -> C.c_void_p:
        """d2r_view for (current training view intrinsics, W x H); cached."""
        self._require()
        s = self.snapshot
        use_lens = bool(self.nerf.render_with_lens_distortion)
        key = (self._view_idx if self._lens_from_view else -1, int(width), int(height), use_lens)
        if key in self._views:
            return self._views[key]
        cam = N.Camera()
        cam.width, cam.height = int(width), int(height)
        if self._lens_from_view:
            v = s.views[self._view_idx]
            rel = v.focal_length / np.float32(v.resolution[s.fov_axis])                 # testbed.cu:456
            sc = np.float32(1.0) - v.principal_point                                     # testbed.cu:464
            lens_mode, lens_params = v.lens_mode, v.lens_params
        else:
            rel, sc = s.snapshot_rel_focal, s.snapshot_screen_center
            lens_mode, lens_params = "Perspective", np.zeros(4, np.float32)
        focal = rel * np.float32((width, height)[s.fov_axis]) * np.float32(s.zoom)       # testbed.cu:4065-4067
        sc = (np.float32(0.5) - sc) * np.float32(s.zoom) + np.float32(0.5)               # testbed.cu:4069-4072
        if not use_lens:
            lens_mode = "Perspective"
        if lens_mode not in ("Perspective", "OpenCV"):
            raise RuntimeError(f"lens mode {lens_mode} is not on the Dream2Real path")
        cam.focal[:] = [float(focal[0]), float(focal[1])]
        cam.screen_center[:] = [float(sc[0]), float(sc[1])]
        cam.lens_mode = 1 if lens_mode == "OpenCV" else 0
        cam.lens_params[:] = [float(x) for x in lens_params[:4]]
        h = C.c_void_p()
        N.check(N.lib().d2r_view_prepare(C.byref(cam), self.device, C.byref(h)), "view_prepare")
        self._views[key] = h
        return h

    def cams_to_ngp(self, cams_nerf: np.ndarray) -> np.ndarray:
        s = self.snapshot
        return np.ascontiguousarray(nerf_matrix_to_ngp(np.asarray(cams_nerf, np.float64)[..., :3, :], s.dataset_scale,
                                                       s.dataset_offset, s.from_mitsuba), dtype=np.float32)

    def render_batch(self, cams, width: int, height: int, ngp_convention: bool = False, want_shade: bool = True,
                     want_depth: bool = True, count_samples: bool = False, want_cost: bool = False):
        """K renders in one launch.  cams: [K,3|4,4] NeRF convention (or NGP if ngp_convention).
        Returns torch float32 CUDA tensors ([K,H,W,4] shade or None, [K,H,W,4] depth or None) and, with want_cost, a
        third one: [K,H,W] step counts (the reference's Cost render mode before the /128)."""
        import torch
        self._require()
        self._sync_min_T()
        cams = np.asarray(cams)
        cams_ngp = np.ascontiguousarray(cams[..., :3, :], np.float32) if ngp_convention else self.cams_to_ngp(cams)
        K = cams_ngp.shape[0]
        view = self.view_handle(width, height)
        dev = torch.device("cuda", self.device)
        shade = torch.empty((K, height, width, 4), dtype=torch.float32, device=dev) if want_shade else None
        depth = torch.empty((K, height, width, 4), dtype=torch.float32, device=dev) if want_depth else None
        cost = torch.empty((K, height, width), dtype=torch.float32, device=dev) if want_cost else None
        ns = torch.zeros(1, dtype=torch.int64, device=dev) if count_samples else None
        with torch.cuda.device(dev):
            N.check(N.lib().d2r_render_ex(self._model, view, cams_ngp.ctypes.data, K, N.f4(self.background_color),
                                          shade.data_ptr() if want_shade else None, depth.data_ptr() if want_depth else None,
                                          cost.data_ptr() if want_cost else None,
                                          ns.data_ptr() if count_samples else None, N.stream_ptr()), "render")
        if count_samples:
            self.last_n_samples = int(ns.item())
        return (shade, depth, cost) if want_cost else (shade, depth)

    def render_composite_batch(self, cams, width: int, height: int, bg_rgba, bg_depth, out_u8=None,
                               ngp_convention: bool = False, count_samples: bool = False, rects_out=None, bg_u8_out=None):
        """K candidate renders composited over a cached background (combined_rendering.py:117-155).
        bg_rgba [H,W,4] f32 cuda, bg_depth [H,W] f32 cuda -> uint8 cuda [K,H,W,3].
        rects_out (int32 cuda [K,4]) / bg_u8_out (uint8 cuda [H,W,3]), optional: the rectangle outside which frame k
        equals the composited background, and that background frame (what ClipVision.preprocess can exploit)."""
        import torch
        self._require()
        self._sync_min_T()
        cams = np.asarray(cams)
        cams_ngp = np.ascontiguousarray(cams[..., :3, :], np.float32) if ngp_convention else self.cams_to_ngp(cams)
        K = cams_ngp.shape[0]
        view = self.view_handle(width, height)
        dev = torch.device("cuda", self.device)
        assert bg_rgba.is_cuda and bg_rgba.dtype == torch.float32 and tuple(bg_rgba.shape) == (height, width, 4) and bg_rgba.is_contiguous()
        assert bg_depth.is_cuda and bg_depth.dtype == torch.float32 and tuple(bg_depth.shape) == (height, width) and bg_depth.is_contiguous()
        if out_u8 is None:
            out_u8 = torch.empty((K, height, width, 3), dtype=torch.uint8, device=dev)
        ns = torch.zeros(1, dtype=torch.int64, device=dev) if count_samples else None
        if rects_out is not None:
            assert rects_out.is_cuda and rects_out.dtype == torch.int32 and tuple(rects_out.shape) == (K, 4) and rects_out.is_contiguous()
        if bg_u8_out is not None:
            assert bg_u8_out.is_cuda and bg_u8_out.dtype == torch.uint8 and tuple(bg_u8_out.shape) == (height, width, 3) and bg_u8_out.is_contiguous()
        with torch.cuda.device(dev):
            N.check(N.lib().d2r_render_composite_ex(self._model, view, cams_ngp.ctypes.data, K, N.f4(self.background_color),
                                                    bg_rgba.data_ptr(), bg_depth.data_ptr(), out_u8.data_ptr(),
                                                    rects_out.data_ptr() if rects_out is not None else None,
                                                    bg_u8_out.data_ptr() if bg_u8_out is not None else None,
                                                    ns.data_ptr() if count_samples else None, N.stream_ptr()), "render_composite")
        if count_samples:
            self.last_n_samples = int(ns.item())
        return out_u8

    # ---- introspection used by the parity tests ----------------------------------------------------
    def occupancy_bitfield(self) -> np.ndarray:
        self._require()
        out = np.empty(128 ** 3 // 8 * 8, np.uint8)
        N.check(N.lib().d2r_model_get_bitfield(self._model, out.ctypes.data, out.size))
        return out

    def occupied_aabb(self) -> np.ndarray:
        self._require()
        out = np.empty(6, np.float32)
        N.check(N.lib().d2r_model_get_occupied_aabb(self._model, out.ctypes.data))
        return out

    def view_dirs(self, width: int, height: int) -> np.ndarray:
        out = np.empty((height, width, 2), np.float32)
        N.check(N.lib().d2r_view_get_dirs(self.view_handle(width, height), out.ctypes.data))
        return out

    # ---- plumbing ----------------------------------------------------------------------------------
    def _sync_min_T(self):
        # nerf.render_min_transmittance may be changed after load (combined_rendering.py:49)
        if self.nerf.render_min_transmittance != self._loaded_min_T:
            N.check(N.lib().d2r_model_set_min_transmittance(self._model, float(self.nerf.render_min_transmittance)))
            self._loaded_min_T = self.nerf.render_min_transmittance

    def _require(self):
        if self.snapshot is None or not self._model:
            raise RuntimeError("no snapshot loaded")

    def _free_native(self, keep_views: bool = False):
        try:
            L = N.lib()
        except Exception:
            return
        if not keep_views:
            for h in self._views.values():
                L.d2r_view_free(h)
            self._views = {}
        if self._model:
            L.d2r_model_free(self._model)
            self._model = C.c_void_p()

    def __del__(self):
        try:
            self._free_native()
        except Exception:
            pass


def free_temporary_memory() -> None:
    """pyngp.free_temporary_memory (python_api.cu:265): release cached allocator blocks."""
    import torch
    if torch.cuda.is_available():
        torch.cuda.empty_cache()

"""Instant-NGP `.ingp` snapshot reader / writer (host side, no GPU work).

A `.ingp` file is zlib/gzip( msgpack( network-config JSON + "snapshot" ) ); see reference
reconstruction/instant-ngp/src/testbed.cu:4691-4755 (save_snapshot) and :4757-4871
(load_snapshot), tiny-cuda-nn trainer.h:275-318 (params_binary / n_params / params_type),
gpu_memory_json.h:36-48 (binary blobs) and vec_json.h:39-59 (matrices are lists of ROWS).

Only the fields the render path consumes are interpreted; everything else is carried
through untouched so a re-written file still loads in the reference.
"""
from __future__ import annotations

import dataclasses
import math
import zlib
from typing import Any, Dict, List, Optional

import msgpack
import numpy as np

NERF_GRIDSIZE = 128                      # nerf_device.cuh:23
NERF_GRID_N_CELLS = NERF_GRIDSIZE ** 3   # nerf_device.cuh:24
NERF_CASCADES = 8                        # nerf_device.cuh:28
SNAPSHOT_FORMAT_VERSION = 1              # testbed.cu:4689


class SnapshotError(RuntimeError):
    """Mirrors the std::runtime_error pybind11 turns into RuntimeError (python_api.cu)."""


@dataclasses.dataclass
class ViewMeta:
    """TrainingImageMetadata subset (nerf_device.cuh:45-59)."""
    focal_length: np.ndarray      # [2] pixels
    principal_point: np.ndarray   # [2] fraction of the image
    resolution: np.ndarray        # [2] int (w, h)
    lens_mode: str                # "Perspective" | "OpenCV" | unsupported others
    lens_params: np.ndarray       # [4] k1 k2 p1 p2


@dataclasses.dataclass
class GridConfig:
    n_levels: int
    n_features_per_level: int
    log2_hashmap_size: int
    base_resolution: int
    per_level_scale: float        # fp32 value, recomputed like testbed.cu:3661-3673
    offsets: np.ndarray           # [n_levels+1] uint32, in ENTRIES (grid.h:693-722)
    scales: np.ndarray            # [n_levels] fp32 grid_scale (common_device.h:856-861)
    resolutions: np.ndarray       # [n_levels] uint32 grid_resolution (common_device.h:863-865)

    @property
    def n_params(self) -> int:
        return int(self.offsets[-1]) * self.n_features_per_level


@dataclasses.dataclass
class Snapshot:
    """Everything the render path needs from one `.ingp` file."""
    config: Dict[str, Any]            # full decoded file (for re-saving)
    grid: GridConfig
    params: np.ndarray                # fp16 [n_params]: density MLP | rgb MLP | hash grid (nerf_network.h:356-372)
    density_mlp: List[np.ndarray]     # fp16 [64,32], [16,64]           (row-major [out,in], fully_fused_mlp.cu:854-863)
    rgb_mlp: List[np.ndarray]         # fp16 [64,32], [64,64], [16,64]
    grid_params: np.ndarray           # fp16 [n_entries, F]
    density_grid: np.ndarray          # fp32 [n_cascades*128^3] (from fp16, testbed.cu:4801-4806)
    aabb_scale: int
    max_cascade: int
    aabb_min: np.ndarray              # m_aabb  (testbed_nerf.cu:2212-2213)
    aabb_max: np.ndarray
    render_aabb_min: np.ndarray       # m_render_aabb (snapshot value, testbed.cu:4843)
    render_aabb_max: np.ndarray
    render_aabb_to_local: np.ndarray  # [3,3]
    cone_angle_constant: float        # testbed_nerf.cu:2228
    background_color: np.ndarray      # [4]
    dataset_scale: float
    dataset_offset: np.ndarray        # [3]
    from_mitsuba: bool
    views: List[ViewMeta]
    fov_axis: int
    zoom: float
    snapshot_screen_center: np.ndarray
    snapshot_rel_focal: np.ndarray
    snapshot_camera: np.ndarray       # [3,4] rows (m_camera)
    exposure: float
    aperture_size: float


# --------------------------------------------------------------------------------------
# grid geometry (tiny-cuda-nn)
# --------------------------------------------------------------------------------------
def per_level_scale_for(aabb_scale: int, base_resolution: int, n_levels: int,
                        desired_resolution: float = 2048.0) -> np.float32:
    """testbed.cu:3661-3673: std::exp(std::log(desired*aabb_scale/base)/(L-1)) in fp32."""
    f = np.float32
    if n_levels <= 1:
        return f(1.0)
    x = f(f(desired_resolution) * f(aabb_scale)) / f(base_resolution)
    return f(np.exp(f(np.log(x)) / f(n_levels - 1), dtype=np.float32))


def grid_geometry(n_levels: int, n_features_per_level: int, log2_hashmap_size: int,
                  base_resolution: int, per_level_scale: float) -> GridConfig:
    """GridEncodingTemplated ctor (grid.h:668-730) with grid_type == Hash."""
    f = np.float32
    log2_pls = f(np.log2(f(per_level_scale)))
    offsets = np.zeros(n_levels + 1, dtype=np.uint32)
    scales = np.zeros(n_levels, dtype=np.float32)
    ress = np.zeros(n_levels, dtype=np.uint32)
    offset = 0
    for lvl in range(n_levels):
        scale = f(f(np.exp2(f(lvl) * log2_pls, dtype=np.float32)) * f(base_resolution)) - f(1.0)
        res = int(math.ceil(float(scale))) + 1
        max_params = (2 ** 32 - 1) // 2
        dense = res ** 3
        params_in_level = max_params if float(res) ** 3 > float(max_params) else dense
        params_in_level = (params_in_level + 7) // 8 * 8
        params_in_level = min(params_in_level, 1 << log2_hashmap_size)
        offsets[lvl] = offset
        offset += params_in_level
        scales[lvl] = scale
        ress[lvl] = res
    offsets[n_levels] = offset
    return GridConfig(n_levels, n_features_per_level, log2_hashmap_size, base_resolution,
                      float(per_level_scale), offsets, scales, ress)


def mlp_shapes(n_in: int, width: int, n_hidden: int, n_out_padded: int):
    """FullyFusedMLP weight matrices, row-major [out,in] (fully_fused_mlp.cu:635-678,854-863)."""
    shapes = [(width, n_in)]
    for _ in range(n_hidden - 1):
        shapes.append((width, width))
    shapes.append((n_out_padded, width))
    return shapes


# --------------------------------------------------------------------------------------
# decode
# --------------------------------------------------------------------------------------
def _decode_bytes(raw: bytes) -> Dict[str, Any]:
    if raw[:2] == b"\x1f\x8b" or raw[:1] == b"\x78":
        raw = zlib.decompress(raw, 15 + 32)  # gzip or zlib header auto-detect (zstr)
    return msgpack.unpackb(raw, raw=False, strict_map_key=False)


def _bin(x) -> bytes:
    if isinstance(x, (bytes, bytearray, memoryview)):
        return bytes(x)
    if isinstance(x, msgpack.ExtType):
        return bytes(x.data)
    if isinstance(x, dict) and "bytes" in x:   # nlohmann json text form
        return bytes(x["bytes"])
    raise SnapshotError("snapshot binary blob has an unexpected encoding")


def _lens_from_json(j: Dict[str, Any]):
    """json_binding.h:66-97: a lens with a "k1" key is OpenCV (or fisheye), no key = Perspective."""
    if "k1" in j:
        if j.get("is_fisheye", False):
            return "OpenCVFisheye", np.array([j["k1"], j["k2"], j["k3"], j["k4"]], np.float32)
        return "OpenCV", np.array([j["k1"], j["k2"], j["p1"], j["p2"]], np.float32)
    if "ftheta_p0" in j:
        return "FTheta", np.zeros(4, np.float32)
    if "latlong" in j:
        return "LatLong", np.zeros(4, np.float32)
    if "equirectangular" in j:
        return "Equirectangular", np.zeros(4, np.float32)
    return "Perspective", np.zeros(4, np.float32)


def load_snapshot(path: str) -> Snapshot:
    with open(path, "rb") as fh:
        raw = fh.read()
    return decode_snapshot(_decode_bytes(raw))


def decode_snapshot(cfg: Dict[str, Any]) -> Snapshot:
    if "snapshot" not in cfg:
        raise SnapshotError("File does not contain a snapshot.")                      # testbed.cu:4876
    snap = cfg["snapshot"]
    if snap.get("version", 0) < SNAPSHOT_FORMAT_VERSION:
        raise SnapshotError("Snapshot uses an old format and can not be loaded.")     # testbed.cu:4759-4761
    if snap.get("mode", "nerf") != "nerf":
        raise SnapshotError("Only NeRF snapshots are on this path.")
    if snap.get("density_grid_size") != NERF_GRIDSIZE:
        raise SnapshotError("Incompatible grid size.")                                # testbed.cu:4776-4778
    if snap.get("params_type", "__half") != "__half":
        raise SnapshotError("params_type must be __half on this path")

    nerf = snap["nerf"]
    ds = nerf.get("dataset", {})
    aabb_scale = int(ds.get("aabb_scale", nerf.get("aabb_scale", 1)))
    if aabb_scale & (aabb_scale - 1):
        raise SnapshotError(f"NeRF dataset's `aabb_scale` must be a power of two, but is {aabb_scale}.")
    if aabb_scale > (1 << (NERF_CASCADES - 1)):
        raise SnapshotError("aabb_scale too large")
    max_cascade = 0
    while (1 << max_cascade) < aabb_scale:
        max_cascade += 1

    enc = cfg["encoding"]
    if "grid" not in enc.get("otype", "").lower():
        raise SnapshotError("only (Hash)Grid position encodings are on this path")
    if enc.get("hash", "CoherentPrime") != "CoherentPrime" or enc.get("interpolation", "Linear") != "Linear":
        raise SnapshotError("only CoherentPrime hash + Linear interpolation are on this path")
    F = int(enc.get("n_features_per_level", 2))
    L = int(enc.get("n_levels", 16))
    log2T = int(enc.get("log2_hashmap_size", 19))
    base_res = int(enc.get("base_resolution", 16))
    pls = enc.get("per_level_scale", 0.0)
    if not pls or pls <= 0:
        pls = per_level_scale_for(aabb_scale, base_res, L)
    grid = grid_geometry(L, F, log2T, base_res, float(pls))
    if F != 4 or L != 8:
        raise SnapshotError("this path supports the Dream2Real NGP config: 8 levels x 4 features")

    net, rgbnet = cfg["network"], cfg["rgb_network"]
    for n in (net, rgbnet):
        if n.get("otype") != "FullyFusedMLP" or int(n.get("n_neurons", 128)) != 64 \
           or n.get("activation", "ReLU") != "ReLU" or n.get("output_activation", "None") != "None":
            raise SnapshotError("this path supports FullyFusedMLP(64, ReLU, None) networks only")
    # what the kernels hard-code beyond the shapes: SH degree 4 on the 3 direction dims and no extra learnable dims (they
    # would widen the rgb-network input, nerf_network.h:105-140).  The density / rgb activations are not part of a snapshot
    # (runtime state, testbed.h:745-746 and testbed_nerf.cu:2145): the path's Exponential / Logistic always apply.
    denc = cfg.get("dir_encoding", {})
    sh = denc["nested"][0] if denc.get("otype") == "Composite" and denc.get("nested") else denc
    if sh.get("otype") != "SphericalHarmonics" or int(sh.get("degree", 4)) != 4:
        raise SnapshotError("this path supports the SphericalHarmonics(degree 4) direction encoding only")
    if int(ds.get("n_extra_learnable_dims", 0)) != 0:
        raise SnapshotError("n_extra_learnable_dims != 0 is not on this path")
    d_shapes = mlp_shapes(L * F, 64, int(net["n_hidden_layers"]), 16)
    c_shapes = mlp_shapes(32, 64, int(rgbnet["n_hidden_layers"]), 16)
    if len(d_shapes) != 2 or len(c_shapes) != 3:
        raise SnapshotError("this path supports density n_hidden_layers=1 and rgb n_hidden_layers=2")

    params = np.frombuffer(_bin(snap["params_binary"]), dtype=np.float16)
    n_mlp = sum(a * b for a, b in d_shapes + c_shapes)
    if params.size != int(snap["n_params"]) or params.size != n_mlp + grid.n_params:
        raise SnapshotError(f"n_params mismatch: file {params.size}, expected {n_mlp + grid.n_params}")
    off = 0
    mats = []
    for (o, i) in d_shapes + c_shapes:
        mats.append(params[off:off + o * i].reshape(o, i))
        off += o * i
    grid_params = params[off:off + grid.n_params].reshape(-1, F)

    dg = np.frombuffer(_bin(snap["density_grid_binary"]), dtype=np.float16).astype(np.float32)
    if dg.size not in (0, NERF_GRID_N_CELLS * (max_cascade + 1)):
        raise SnapshotError("Incompatible number of grid cascades.")                  # testbed.cu:4811-4814

    half = 0.5 * min(1 << (NERF_CASCADES - 1), aabb_scale)
    aabb_min = np.full(3, 0.5 - half, np.float32)
    aabb_max = np.full(3, 0.5 + half, np.float32)
    ra = snap.get("render_aabb", {"min": aabb_min.tolist(), "max": aabb_max.tolist()})
    r2l = np.array(snap.get("render_aabb_to_local", np.eye(3).tolist()), np.float32)

    views = []
    for md in ds.get("metadata", []):
        mode, lp = _lens_from_json(md.get("lens", {}))
        views.append(ViewMeta(np.array(md["focal_length"], np.float32), np.array(md["principal_point"], np.float32),
                              np.array(md["resolution"], np.int32), mode, lp))
    cam = snap.get("camera", {})
    return Snapshot(
        config=cfg, grid=grid, params=params, density_mlp=mats[:2], rgb_mlp=mats[2:], grid_params=grid_params,
        density_grid=dg, aabb_scale=aabb_scale, max_cascade=max_cascade, aabb_min=aabb_min, aabb_max=aabb_max,
        render_aabb_min=np.array(ra["min"], np.float32), render_aabb_max=np.array(ra["max"], np.float32),
        render_aabb_to_local=r2l,
        cone_angle_constant=0.0 if aabb_scale <= 1 else 1.0 / 256.0,
        background_color=np.array(snap.get("background_color", [0, 0, 0, 1]), np.float32),
        dataset_scale=float(ds.get("scale", 0.33)), dataset_offset=np.array(ds.get("offset", [0.5, 0.5, 0.5]), np.float32),
        from_mitsuba=bool(ds.get("from_mitsuba", False)), views=views,
        fov_axis=int(cam.get("fov_axis", 1)), zoom=float(cam.get("zoom", 1.0)),
        snapshot_screen_center=np.array(cam.get("screen_center", [0.5, 0.5]), np.float32),
        snapshot_rel_focal=np.array(cam.get("relative_focal_length", [1.0, 1.0]), np.float32),
        snapshot_camera=np.array(cam.get("matrix", [[1, 0, 0, 0.5], [0, -1, 0, 0.5], [0, 0, -1, 0.5]]), np.float32),
        exposure=float(snap.get("exposure", 0.0)), aperture_size=float(cam.get("aperture_size", 0.0)),
    )


# --------------------------------------------------------------------------------------
# encode (used by the synthetic-scene generator and the fixture re-packer)
# --------------------------------------------------------------------------------------
BASE_NETWORK_CONFIG = {
    # reference reconstruction/instant-ngp/configs/nerf/base.json (the only config Dream2Real uses)
    "loss": {"otype": "Huber"},
    "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {
        "otype": "ExponentialDecay", "decay_start": 20000, "decay_interval": 10000, "decay_base": 0.33,
        "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
    "encoding": {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 4, "log2_hashmap_size": 19, "base_resolution": 16},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1},
    "dir_encoding": {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4}, {"otype": "Identity"}]},
    "rgb_network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
}


def _adam_stub(n: int) -> Dict[str, Any]:
    z = [0.0] * n
    return {"beta1": 0.9, "beta2": 0.99, "epsilon": 1e-8, "first_moment": z, "iter": 0,
            "learning_rate": 1e-4, "second_moment": z, "variable": z}


def build_snapshot_config(params_f16: np.ndarray, density_grid_f16: np.ndarray, *, aabb_scale: int,
                          log2_hashmap_size: int, views: List[Dict[str, Any]], xforms: List[np.ndarray],
                          scale: float, offset, background_color=(0.0, 0.0, 0.0, 0.0)) -> Dict[str, Any]:
    """Build a dict the reference's Testbed::load_snapshot accepts (testbed.cu:4757-4871)."""
    import copy
    cfg = copy.deepcopy(BASE_NETWORK_CONFIG)
    cfg["encoding"]["log2_hashmap_size"] = int(log2_hashmap_size)
    half = 0.5 * aabb_scale
    n = len(views)
    inf = float("inf")
    dataset = {
        "aabb_scale": int(aabb_scale), "envmap_resolution": [0, 0], "from_mitsuba": False, "is_hdr": False,
        "metadata": [{
            "focal_length": [float(v["fl_x"]), float(v["fl_y"])],
            "lens": ({"is_fisheye": False, "k1": float(v.get("k1", 0)), "k2": float(v.get("k2", 0)),
                      "p1": float(v.get("p1", 0)), "p2": float(v.get("p2", 0))} if "k1" in v else {}),
            "principal_point": [float(v["cx"]) / float(v["w"]), float(v["cy"]) / float(v["h"])],
            "resolution": [int(v["w"]), int(v["h"])], "rolling_shutter": [0.0, 0.0, 0.0, 0.0],
        } for v in views],
        "n_extra_learnable_dims": 0, "n_images": n, "offset": [float(o) for o in offset],
        "paths": [f"images/{i:04d}.png" for i in range(n)],
        "render_aabb": {"max": [-inf] * 3, "min": [inf] * 3},
        "render_aabb_to_local": np.eye(3).tolist(), "scale": float(scale), "up": [0.0, 1.0, 0.0],
        "wants_importance_sampling": True,
        "xforms": [{"start": np.asarray(x, np.float64)[:3].tolist(), "end": np.asarray(x, np.float64)[:3].tolist()} for x in xforms],
    }
    cfg["snapshot"] = {
        "aabb": {"min": [0.5 - half] * 3, "max": [0.5 + half] * 3},
        "background_color": [float(c) for c in background_color], "bounding_radius": 1.0,
        "camera": {"aperture_size": 0.0, "autofocus": False, "autofocus_depth": 0.0, "autofocus_target": [0.5, 0.5, 0.5],
                   "fov_axis": 1, "matrix": [[1.0, 0.0, 0.0, 0.5], [0.0, -1.0, 0.0, 0.5], [0.0, 0.0, -1.0, 2.0]],
                   "relative_focal_length": [1.0, 1.0], "scale": 1.5, "screen_center": [0.5, 0.5], "zoom": 1.0},
        "density_grid_binary": np.ascontiguousarray(density_grid_f16, np.float16).tobytes(),
        "density_grid_size": NERF_GRIDSIZE, "exposure": 0.0, "loss": 0.0, "mode": "nerf",
        "n_params": int(params_f16.size),
        "nerf": {"aabb_scale": int(aabb_scale),
                 "cam_pos_offset": [_adam_stub(3) for _ in range(n)], "cam_rot_offset": [_adam_stub(3) for _ in range(n)],
                 "extra_dims_opt": [_adam_stub(0) for _ in range(n)], "dataset": dataset,
                 "rgb": {"measured_batch_size": 0, "measured_batch_size_before_compaction": 0, "rays_per_batch": 4096}},
        "params_binary": np.ascontiguousarray(params_f16, np.float16).tobytes(), "params_type": "__half",
        "render_aabb": {"min": [0.5 - half] * 3, "max": [0.5 + half] * 3},
        "render_aabb_to_local": np.eye(3).tolist(),
        "sun_dir": [0.5773502588272095] * 3, "training_step": 0, "up_dir": [0.0, 1.0, 0.0],
        "version": SNAPSHOT_FORMAT_VERSION,
    }
    return cfg


def save_snapshot(path: str, cfg: Dict[str, Any], compress_level: int = 6) -> None:
    """gzip(msgpack(cfg)) like zstr::ostream + json::to_msgpack (testbed.cu:4743-4751)."""
    payload = msgpack.packb(cfg, use_bin_type=True)
    co = zlib.compressobj(compress_level, zlib.DEFLATED, 15 + 16)
    with open(path, "wb") as fh:
        fh.write(co.compress(payload) + co.flush())

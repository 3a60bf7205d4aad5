"""CPU oracle: numpy restatement of the Instant-NGP NeRF *render* path Dream2Real calls.

TEST INFRASTRUCTURE ONLY.  Nothing under dream2real_b200/ imports this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.

Pinning: (1) the known-answer vectors host-compiled from the reference's own NGP_HOST_DEVICE
headers (tests/golden/kat_host.json, kat_cam.json);  (2) full renders of the
REAL reference binary (pyngp built from /root/reference/reconstruction/instant-ngp) executed on a
B200 by tests/golden/make_golden_pyngp.py -> tests/golden/*.npz.  See tests/test_oracle_*.py.

All paths below are relative to /root/reference/reconstruction/instant-ngp ("TCNN" =
dependencies/tiny-cuda-nn).  The reference is compiled with --use_fast_math and runs its
MLPs on fp16 tensor cores, so float results agree to tolerance, integer results bit-exactly.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import numpy as np

f32 = np.float32
u32 = np.uint32

# ----------------------------------------------------------------------------------------------
# constants (include/neural-graphics-primitives/nerf_device.cuh:23-42, common_device.cuh:32)
# ----------------------------------------------------------------------------------------------
NERF_GRIDSIZE = 128
NERF_GRID_N_CELLS = 128 ** 3
NERF_CASCADES = 8
NERF_STEPS = 1024
SQRT3 = f32(1.73205080757)
STEPSIZE = f32(SQRT3 / f32(NERF_STEPS))
MIN_CONE_STEPSIZE = STEPSIZE
MAX_CONE_STEPSIZE = f32(f32(f32(STEPSIZE * f32(1 << (NERF_CASCADES - 1))) * f32(NERF_STEPS)) / f32(NERF_GRIDSIZE))
MAX_DEPTH = f32(16384.0)
NERF_MIN_OPTICAL_THICKNESS = f32(0.01)
MARCH_ITER = 10000  # src/testbed_nerf.cu:59


# ----------------------------------------------------------------------------------------------
# integer helpers  (bit-exact)
# ----------------------------------------------------------------------------------------------
def expand_bits(v):
    """TCNN common_device.h:761-767."""
    v = np.asarray(v, dtype=u32)
    v = (v * u32(0x00010001)) & u32(0xFF0000FF)
    v = (v * u32(0x00000101)) & u32(0x0F00F00F)
    v = (v * u32(0x00000011)) & u32(0xC30C30C3)
    v = (v * u32(0x00000005)) & u32(0x49249249)
    return v


def morton3D(x, y, z):
    """TCNN common_device.h:771-776."""
    return expand_bits(x) | (expand_bits(y) << u32(1)) | (expand_bits(z) << u32(2))


def morton3D_invert(x):
    """TCNN common_device.h:778-785."""
    x = np.asarray(x, dtype=u32) & u32(0x49249249)
    x = (x | (x >> u32(2))) & u32(0xC30C30C3)
    x = (x | (x >> u32(4))) & u32(0x0F00F00F)
    x = (x | (x >> u32(8))) & u32(0xFF0000FF)
    x = (x | (x >> u32(16))) & u32(0x0000FFFF)
    return x


def reverse_bits(x):
    """random_val.cuh:219-226."""
    x = np.asarray(x, dtype=u32)
    x = ((x & u32(0xAAAAAAAA)) >> u32(1)) | ((x & u32(0x55555555)) << u32(1))
    x = ((x & u32(0xCCCCCCCC)) >> u32(2)) | ((x & u32(0x33333333)) << u32(2))
    x = ((x & u32(0xF0F0F0F0)) >> u32(4)) | ((x & u32(0x0F0F0F0F)) << u32(4))
    x = ((x & u32(0xFF00FF00)) >> u32(8)) | ((x & u32(0x00FF00FF)) << u32(8))
    return (x >> u32(16)) | (x << u32(16))


def laine_karras_permutation(x, seed):
    """random_val.cuh:228-235."""
    x = np.asarray(x, dtype=u32)
    with np.errstate(over="ignore"):
        x = x + u32(seed) if np.isscalar(seed) else x + np.asarray(seed, u32)
        x = x ^ (x * u32(0x6C50B47C))
        x = x ^ (x * u32(0xB82F1E52))
        x = x ^ (x * u32(0xC7AFE638))
        x = x ^ (x * u32(0x8D22F6E6))
    return x


def nested_uniform_scramble_base2(x, seed):
    """random_val.cuh:237-242."""
    return reverse_bits(laine_karras_permutation(reverse_bits(x), seed))


def hash_combine(seed, v):
    """random_val.cuh:215-217."""
    seed = np.asarray(seed, u32)
    with np.errstate(over="ignore"):
        return seed ^ (u32(v) + (seed << u32(6)) + (seed >> u32(2)))


_SOBOL_DIM1 = np.array([
    0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000, 0xcc000000, 0xaa000000, 0xff000000,
    0x80800000, 0xc0c00000, 0xa0a00000, 0xf0f00000, 0x88880000, 0xcccc0000, 0xaaaa0000, 0xffff0000,
    0x80008000, 0xc000c000, 0xa000a000, 0xf000f000, 0x88008800, 0xcc00cc00, 0xaa00aa00, 0xff00ff00,
    0x80808080, 0xc0c0c0c0, 0xa0a0a0a0, 0xf0f0f0f0, 0x88888888, 0xcccccccc, 0xaaaaaaaa, 0xffffffff], dtype=u32)


def sobol(index, dim):
    """random_val.cuh:162-209 (dims 0 and 1 are all this path touches)."""
    index = np.asarray(index, u32)
    if dim == 0:
        return reverse_bits(index)  # direction numbers of dim 0 are 0x80000000 >> bit
    assert dim == 1
    X = np.zeros_like(index)
    for bit in range(32):
        X = X ^ (((index >> u32(bit)) & u32(1)) * _SOBOL_DIM1[bit])
    return X


def ld_random_val(index, seed, dim=0):
    """random_val.cuh:287-291."""
    seed = np.asarray(seed, u32)
    idx = nested_uniform_scramble_base2(np.full_like(seed, index), seed)
    x = nested_uniform_scramble_base2(sobol(idx, dim), hash_combine(seed, dim))
    return x.astype(f32) * f32(1.0 / (1 << 32))


def ld_random_val_2d(index, seed):
    """random_val.cuh:280-284 (shuffled_scrambled_sobol2d)."""
    seed = np.asarray([seed], u32)
    idx = nested_uniform_scramble_base2(np.full_like(seed, index), seed)
    out = []
    for i in range(2):
        x = nested_uniform_scramble_base2(sobol(idx, i), hash_combine(seed, i))
        out.append(x.astype(f32)[0] * f32(1.0 / (1 << 32)))
    return np.array(out, f32)


def ld_random_pixel_offset(spp):
    """random_val.cuh:320-325."""
    o = f32(0.5) - ld_random_val_2d(0, 0xdeadbeef) + ld_random_val_2d(spp, 0xdeadbeef)
    return (o - np.floor(o)).astype(f32)


# ----------------------------------------------------------------------------------------------
# stepping space  (nerf_device.cuh:378-440)
# ----------------------------------------------------------------------------------------------
def _step_consts(cone):
    log1p_c = f32(np.log(f32(1.0) + f32(cone)))
    a = f32((f32(np.log(MIN_CONE_STEPSIZE)) - f32(np.log(log1p_c))) / log1p_c)
    b = f32((f32(np.log(MAX_CONE_STEPSIZE)) - f32(np.log(log1p_c))) / log1p_c)
    at = f32(np.exp(f32(a * log1p_c)))
    bt = f32(np.exp(f32(b * log1p_c)))
    return log1p_c, a, b, at, bt


def to_stepping_space(t, cone):
    t = np.asarray(t, f32)
    if cone <= 1e-5:
        return (t / MIN_CONE_STEPSIZE).astype(f32)
    l, a, b, at, bt = _step_consts(cone)
    with np.errstate(invalid="ignore", divide="ignore"):
        mid = (np.log(t) / l).astype(f32)
    lo = ((t - at) / MIN_CONE_STEPSIZE + a).astype(f32)
    hi = ((t - bt) / MAX_CONE_STEPSIZE + b).astype(f32)
    return np.where(t <= at, lo, np.where(t <= bt, mid, hi)).astype(f32)


def from_stepping_space(n, cone):
    n = np.asarray(n, f32)
    if cone <= 1e-5:
        return (n * MIN_CONE_STEPSIZE).astype(f32)
    l, a, b, at, bt = _step_consts(cone)
    with np.errstate(over="ignore"):
        mid = np.exp((n * l).astype(f32)).astype(f32)
    lo = ((n - a) * MIN_CONE_STEPSIZE + at).astype(f32)
    hi = ((n - b) * MAX_CONE_STEPSIZE + bt).astype(f32)
    return np.where(n <= a, lo, np.where(n <= b, mid, hi)).astype(f32)


def advance_n_steps(t, cone, n):
    return from_stepping_space((to_stepping_space(t, cone) + np.asarray(n, f32)).astype(f32), cone)


def calc_dt(t, cone):
    t = np.asarray(t, f32)
    return (advance_n_steps(t, cone, f32(1.0)) - t).astype(f32)


def warp_dt(dt):
    """nerf_device.cuh:306-309."""
    mx = f32(MIN_CONE_STEPSIZE * f32(1 << (NERF_CASCADES - 1)))
    return ((np.asarray(dt, f32) - MIN_CONE_STEPSIZE) / f32(mx - MIN_CONE_STEPSIZE)).astype(f32)


def unwarp_dt(dt):
    """nerf_device.cuh:311-314."""
    mx = f32(MIN_CONE_STEPSIZE * f32(1 << (NERF_CASCADES - 1)))
    return (np.asarray(dt, f32) * f32(mx - MIN_CONE_STEPSIZE) + MIN_CONE_STEPSIZE).astype(f32)


def mip_from_pos(pos, max_cascade=NERF_CASCADES - 1):
    """nerf_device.cuh:442-447."""
    maxval = np.max(np.abs(np.asarray(pos, f32) - f32(0.5)), axis=-1)
    _, e = np.frexp(maxval)
    return np.clip(e + 1, 0, max_cascade).astype(np.int32)


def cascaded_grid_idx_at(pos, mip):
    """nerf_device.cuh:316-328; returns uint32 with 0xFFFFFFFF = outside."""
    pos = np.asarray(pos, f32)
    mip = np.asarray(mip, np.int32)
    mip_scale = np.ldexp(f32(1.0), -mip).astype(f32)[..., None]
    p = ((pos - f32(0.5)) * mip_scale + f32(0.5)).astype(f32)
    with np.errstate(invalid="ignore"):
        i = np.trunc((p * f32(NERF_GRIDSIZE)).astype(f32)).astype(np.int64)   # C float->int truncation
    bad = np.any((i < 0) | (i >= NERF_GRIDSIZE), axis=-1)
    i = np.clip(i, 0, NERF_GRIDSIZE - 1).astype(u32)
    idx = morton3D(i[..., 0], i[..., 1], i[..., 2])
    return np.where(bad, u32(0xFFFFFFFF), idx)


def density_grid_occupied_at(pos, bitfield, mip):
    """nerf_device.cuh:334-340."""
    idx = cascaded_grid_idx_at(pos, mip)
    ok = idx != u32(0xFFFFFFFF)
    safe = np.where(ok, idx, 0).astype(np.int64)
    byte = bitfield[safe // 8 + (np.asarray(mip, np.int64) * NERF_GRID_N_CELLS) // 8]
    return ok & ((byte >> (safe % 8).astype(np.uint8)) & 1).astype(bool)


def distance_to_next_voxel(pos, d, idir, res):
    """nerf_device.cuh:359-367; res broadcast per ray."""
    res = np.asarray(res, f32)[..., None]
    p = (res * (pos - f32(0.5))).astype(f32)
    sgn = np.copysign(f32(1.0), d).astype(f32)
    with np.errstate(invalid="ignore", over="ignore"):
        t3 = ((np.floor((p + f32(0.5) + f32(0.5) * sgn).astype(f32)) - p) * idir).astype(f32)
    t = np.min(t3, axis=-1)
    return np.maximum((t / res[..., 0]).astype(f32), f32(0.0))


def advance_to_next_voxel(t, cone, pos, d, idir, mip):
    """nerf_device.cuh:430-440."""
    res = np.ldexp(f32(NERF_GRIDSIZE), -np.asarray(mip, np.int32)).astype(f32)
    t_target = (t + distance_to_next_voxel(pos, d, idir, res)).astype(f32)
    ts = to_stepping_space(t, cone)
    tt = to_stepping_space(t_target, cone)
    return from_stepping_space((ts + np.ceil(np.maximum((tt - ts).astype(f32), f32(0.5)))).astype(f32), cone)


def skip_to_occupied(t, cone, o, d, idir, bitfield, max_mip, aabb_min, aabb_max, r2l=None):
    """if_unoccupied_advance_to_next_occupied_voxel (nerf_device.cuh:462-494), min_mip = 0.

    Vectorised over rays; returns t (MAX_DEPTH where the ray left the render aabb)."""
    t = np.array(t, f32, copy=True)
    active = np.ones(t.shape, bool)
    while active.any():
        ia = np.nonzero(active)[0]
        ta = t[ia]
        pos = (o[ia] + ta[:, None] * d[ia]).astype(f32)
        lp = pos if r2l is None else (pos @ r2l.T).astype(f32)
        out = (ta >= MAX_DEPTH) | ~np.all((lp >= aabb_min) & (lp <= aabb_max), axis=-1)
        t[ia[out]] = MAX_DEPTH
        active[ia[out]] = False
        keep = ~out
        ia, ta, pos = ia[keep], ta[keep], pos[keep]
        if ia.size == 0:
            break
        mip = np.clip(mip_from_pos(pos), 0, max_mip)
        occ = density_grid_occupied_at(pos, bitfield, mip)
        active[ia[occ]] = False
        ia, ta, pos, mip = ia[~occ], ta[~occ], pos[~occ], mip[~occ]
        if ia.size == 0:
            break
        # climb to the largest empty voxel
        grow = np.ones(ia.shape, bool)
        while True:
            grow &= mip < max_mip
            if not grow.any():
                break
            g = np.nonzero(grow)[0]
            nocc = ~density_grid_occupied_at(pos[g], bitfield, mip[g] + 1)
            mip[g[nocc]] += 1
            grow[g[~nocc]] = False
        t[ia] = advance_to_next_voxel(ta, cone, pos, d[ia], idir[ia], mip)
    return t


# ----------------------------------------------------------------------------------------------
# colour helpers (common_device.cuh:34-64)
# ----------------------------------------------------------------------------------------------
def srgb_to_linear(x):
    x = np.asarray(x, f32)
    with np.errstate(invalid="ignore"):
        hi = np.power(((x + f32(0.055)) / f32(1.055)).astype(f32), f32(2.4)).astype(f32)
    return np.where(x <= f32(0.04045), (x / f32(12.92)).astype(f32), hi).astype(f32)


def linear_to_srgb_cpp(x):
    x = np.asarray(x, f32)
    with np.errstate(invalid="ignore"):
        hi = (f32(1.055) * np.power(x, f32(0.41666)).astype(f32) - f32(0.055)).astype(f32)
    return np.where(x < f32(0.0031308), (f32(12.92) * x).astype(f32), hi).astype(f32)


def logistic(x):
    """TCNN common_device.h:42-44."""
    x = np.asarray(x, f32)
    with np.errstate(over="ignore"):
        return (f32(1.0) / (f32(1.0) + np.exp(-x).astype(f32))).astype(f32)


# ----------------------------------------------------------------------------------------------
# occupancy bitfield at load (src/testbed_nerf.cu:284-331, 2355-2373)
# ----------------------------------------------------------------------------------------------
def build_bitfield(density_grid: np.ndarray, max_cascade: int) -> Tuple[np.ndarray, float]:
    g = np.asarray(density_grid, f32)
    n = NERF_GRID_N_CELLS
    mean = f32(np.sum(np.maximum(g[:n], f32(0.0)).astype(np.float64) / n))   # reduce_sum over cascade 0 only
    thresh = min(NERF_MIN_OPTICAL_THICKNESS, mean)
    bits = np.zeros(n // 8 * NERF_CASCADES, np.uint8)
    nz = n // 8 * (max_cascade + 1)
    occ = (g[: nz * 8] > thresh).reshape(-1, 8)
    bits[:nz] = np.packbits(occ, axis=1, bitorder="little")[:, 0]
    for level in range(1, NERF_CASCADES):
        prev = bits[(level - 1) * (n // 8): level * (n // 8)]
        nxt = bits[level * (n // 8): (level + 1) * (n // 8)]
        i = np.arange(n // 64, dtype=u32)
        b = np.packbits((prev.reshape(-1, 8) > 0), axis=1, bitorder="little")[:, 0]
        x = morton3D_invert(i >> u32(0)) + u32(NERF_GRIDSIZE // 8)
        y = morton3D_invert(i >> u32(1)) + u32(NERF_GRIDSIZE // 8)
        z = morton3D_invert(i >> u32(2)) + u32(NERF_GRIDSIZE // 8)
        np.bitwise_or.at(nxt, morton3D(x, y, z).astype(np.int64), b)
    return bits, float(thresh)


# ----------------------------------------------------------------------------------------------
# camera (src/testbed.cu:401-403,453-468,4065-4072; nerf_loader.h:101-121)
# ----------------------------------------------------------------------------------------------
def nerf_matrix_to_ngp(m34, scale, offset, from_mitsuba=False):
    m = np.array(m34, f32, copy=True)[:3, :4]
    m[:, 1] *= f32(-1.0)
    m[:, 2] *= f32(-1.0)
    m[:, 3] = (m[:, 3] * f32(scale) + np.asarray(offset, f32)).astype(f32)
    if from_mitsuba:
        m[:, 0] *= f32(-1.0)
        m[:, 2] *= f32(-1.0)
        return m
    return m[[1, 2, 0], :]       # cycle axes xyz <- yzx (rows)


@dataclasses.dataclass
class ViewSetup:
    """State after Testbed::set_camera_to_training_view(v) for a W x H render."""
    W: int
    H: int
    focal: np.ndarray          # [2] pixels  = rel_focal * res[fov_axis] * zoom
    screen_center: np.ndarray  # [2]
    lens_mode: str
    lens_params: np.ndarray


def view_setup(snap, view: int, W: int, H: int) -> ViewSetup:
    v = snap.views[view]
    rel = (v.focal_length / f32(v.resolution[snap.fov_axis])).astype(f32)          # testbed.cu:456
    focal = (rel * f32((W, H)[snap.fov_axis]) * f32(snap.zoom)).astype(f32)        # testbed.cu:4065-4067
    sc = (f32(1.0) - v.principal_point).astype(f32)                                # testbed.cu:464
    sc = ((f32(0.5) - sc) * f32(snap.zoom) + f32(0.5)).astype(f32)                 # testbed.cu:4069-4072
    return ViewSetup(W, H, focal, sc, v.lens_mode, v.lens_params.astype(f32))


def _opencv_delta(p, u, v):
    """common_device.cuh:249-262."""
    k1, k2, p1, p2 = (f32(x) for x in p[:4])
    u2 = u * u
    uv = u * v
    v2 = v * v
    r2 = u2 + v2
    radial = k1 * r2 + k2 * r2 * r2
    du = u * radial + f32(2) * p1 * uv + p2 * (r2 + f32(2) * u2)
    dv = v * radial + f32(2) * p2 * uv + p1 * (r2 + f32(2) * v2)
    return du.astype(f32), dv.astype(f32)


def iterative_opencv_lens_undistortion(params, u, v):
    """common_device.cuh:289-333, vectorised (each ray stops on its own criterion)."""
    u = np.array(u, f32, copy=True)
    v = np.array(v, f32, copy=True)
    x0u, x0v = u.copy(), v.copy()
    eps = np.finfo(f32).eps
    active = np.ones(u.shape, bool)
    for _ in range(100):
        if not active.any():
            break
        a = active
        xu, xv = u[a], v[a]
        s0 = np.maximum(eps, np.abs(f32(1e-6) * xu)).astype(f32)
        s1 = np.maximum(eps, np.abs(f32(1e-6) * xv)).astype(f32)
        dxu, dxv = _opencv_delta(params, xu, xv)
        b0u, b0v = _opencv_delta(params, (xu - s0).astype(f32), xv)
        f0u, f0v = _opencv_delta(params, (xu + s0).astype(f32), xv)
        b1u, b1v = _opencv_delta(params, xu, (xv - s1).astype(f32))
        f1u, f1v = _opencv_delta(params, xu, (xv + s1).astype(f32))
        J00 = (f32(1) + (f0u - b0u) / (f32(2) * s0)).astype(f32)   # d(du)/du
        J10 = ((f1u - b1u) / (f32(2) * s1)).astype(f32)            # d(du)/dv  (column 1, row 0)
        J01 = ((f0v - b0v) / (f32(2) * s0)).astype(f32)            # d(dv)/du
        J11 = (f32(1) + (f1v - b1v) / (f32(2) * s1)).astype(f32)
        ru = (xu + dxu - x0u[a]).astype(f32)
        rv = (xv + dxv - x0v[a]).astype(f32)
        det = (J00 * J11 - J10 * J01).astype(f32)
        su = ((J11 * ru - J10 * rv) / det).astype(f32)
        sv = ((-J01 * ru + J00 * rv) / det).astype(f32)
        u[a] = (xu - su).astype(f32)
        v[a] = (xv - sv).astype(f32)
        done = (su * su + sv * sv) < f32(1e-10)
        idx = np.nonzero(a)[0]
        active[idx[done]] = False
    return u, v


def camera_plane_dirs(vs: ViewSetup) -> np.ndarray:
    """Per-pixel camera-space direction (x, y, 1) before rotation: uv_to_ray (common_device.cuh:393-431)
    with snap_to_pixel_centers off and sample_index 0 -> ld_random_pixel_offset(0) == (0.5, 0.5)."""
    off = ld_random_pixel_offset(0)
    xs = (np.arange(vs.W, dtype=f32) + off[0]) / f32(vs.W)
    ys = (np.arange(vs.H, dtype=f32) + off[1]) / f32(vs.H)
    u = np.broadcast_to(xs[None, :], (vs.H, vs.W)).astype(f32)
    v = np.broadcast_to(ys[:, None], (vs.H, vs.W)).astype(f32)
    dx = ((u - vs.screen_center[0]) * f32(vs.W) / vs.focal[0]).astype(f32).reshape(-1)
    dy = ((v - vs.screen_center[1]) * f32(vs.H) / vs.focal[1]).astype(f32).reshape(-1)
    if vs.lens_mode == "OpenCV":
        dx, dy = iterative_opencv_lens_undistortion(vs.lens_params, dx, dy)
    elif vs.lens_mode != "Perspective":
        raise NotImplementedError(f"lens mode {vs.lens_mode} is not on the Dream2Real path")
    return np.stack([dx, dy, np.ones_like(dx)], axis=-1).astype(f32)


def ray_aabb(o, d, mn, mx):
    """bounding_box.cuh:163-213 -> tmin (FLT_MAX when missed)."""
    FMAX = np.finfo(f32).max
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t0 = ((mn - o) / d).astype(f32)
        t1 = ((mx - o) / d).astype(f32)
    lo = np.minimum(t0, t1)
    hi = np.maximum(t0, t1)
    tmin, tmax = lo[:, 0].copy(), hi[:, 0].copy()
    miss = np.zeros(tmin.shape, bool)
    for ax in (1, 2):
        miss |= (tmin > hi[:, ax]) | (lo[:, ax] > tmax)
        tmin = np.where(lo[:, ax] > tmin, lo[:, ax], tmin)
        tmax = np.where(hi[:, ax] < tmax, hi[:, ax], tmax)
    return np.where(miss, FMAX, tmin).astype(f32)


# ----------------------------------------------------------------------------------------------
# network  (TCNN grid.h:47-165, common_device.h:340-365,631-718,826-870; nerf_network.h:105-140)
# ----------------------------------------------------------------------------------------------
def grid_index(grid, level, pg):
    """grid_index<3, CoherentPrime> (TCNN common_device.h:697-713); pg uint32 [n,3]."""
    hashmap_size = int(grid.offsets[level + 1] - grid.offsets[level])
    res = int(grid.resolutions[level])
    stride = 1
    index = np.zeros(pg.shape[0], u32)
    with np.errstate(over="ignore"):
        for dim in range(3):
            if stride > hashmap_size:
                break
            index = index + pg[:, dim] * u32(stride & 0xFFFFFFFF)
            stride *= res
        if hashmap_size < stride:
            index = pg[:, 0] * u32(1) ^ pg[:, 1] * u32(2654435761) ^ pg[:, 2] * u32(805459861)
    return (index % u32(hashmap_size)).astype(np.int64)


def hash_encode(snap, x01: np.ndarray) -> np.ndarray:
    """kernel_grid<__half,3,4,CoherentPrime>: fp32 positions in [0,1]^3 -> fp16 [n, 32].

    The trilinear sum is an fp16 fma chain with the weight rounded to fp16 first (grid.h:162)."""
    g = snap.grid
    n = x01.shape[0]
    out = np.zeros((n, g.n_levels * g.n_features_per_level), np.float16)
    for lvl in range(g.n_levels):
        scale = f32(g.scales[lvl])
        table = snap.grid_params[int(g.offsets[lvl]): int(g.offsets[lvl + 1])]
        pos = np.float32(np.float64(scale) * x01.astype(np.float64) + 0.5)   # fmaf(scale, x, 0.5f): one rounding
        fl = np.floor(pos)
        pg = fl.astype(np.int64).astype(u32)
        w = (pos - fl).astype(f32)
        res = np.zeros((n, g.n_features_per_level), np.float16)
        for idx in range(8):
            wt = np.ones(n, f32)
            pl = pg.copy()
            for dim in range(3):
                if idx & (1 << dim):
                    wt = (wt * w[:, dim]).astype(f32)
                    pl[:, dim] = pg[:, dim] + u32(1)
                else:
                    wt = (wt * (f32(1) - w[:, dim])).astype(f32)
            val = table[grid_index(g, lvl, pl)]
            wh = wt.astype(np.float16)
            res = (wh.astype(np.float64)[:, None] * val.astype(np.float64) + res.astype(np.float64)).astype(np.float16)
        out[:, lvl * g.n_features_per_level:(lvl + 1) * g.n_features_per_level] = res
    return out


def sh_encode_deg4(d: np.ndarray) -> np.ndarray:
    """sh_enc degree 4 (TCNN common_device.h:340-365) of the *unwarped* direction; fp16 [n,16]."""
    d = np.asarray(d, f32)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = np.empty((d.shape[0], 16), f32)
    o[:, 0] = f32(0.28209479177387814)
    o[:, 1] = f32(-0.48860251190291987) * y
    o[:, 2] = f32(0.48860251190291987) * z
    o[:, 3] = f32(-0.48860251190291987) * x
    o[:, 4] = f32(1.0925484305920792) * xy
    o[:, 5] = f32(-1.0925484305920792) * yz
    o[:, 6] = f32(0.94617469575755997) * z2 - f32(0.31539156525251999)
    o[:, 7] = f32(-1.0925484305920792) * xz
    o[:, 8] = f32(0.54627421529603959) * x2 - f32(0.54627421529603959) * y2
    o[:, 9] = f32(0.59004358992664352) * y * (f32(-3.0) * x2 + y2)
    o[:, 10] = f32(2.8906114426405538) * xy * z
    o[:, 11] = f32(0.45704579946446572) * y * (f32(1.0) - f32(5.0) * z2)
    o[:, 12] = f32(0.3731763325901154) * z * (f32(5.0) * z2 - f32(3.0))
    o[:, 13] = f32(0.45704579946446572) * x * (f32(1.0) - f32(5.0) * z2)
    o[:, 14] = f32(1.4453057213202769) * z * (x2 - y2)
    o[:, 15] = f32(0.59004358992664352) * x * (-x2 + f32(3.0) * y2)
    return o.astype(np.float16)


def _mlp_layer(a16: np.ndarray, w16: np.ndarray, relu: bool, accum: str) -> np.ndarray:
    """One FullyFusedMLP layer on fp16 tensor cores (TCNN fully_fused_mlp.cu:47-129,315-476).

    accum='fp32'    : exact dot product rounded once to fp16.
    accum='fp16_k16': wmma::accumulator<__half> -- partial sums rounded to fp16 after every
                      16-wide k block (the m16n16k16 fragment granularity)."""
    a = a16.astype(np.float32)
    w = w16.astype(np.float32)
    if accum == "fp32":
        r = (a.astype(np.float64) @ w.T.astype(np.float64)).astype(np.float16)
    else:
        r = np.zeros((a.shape[0], w.shape[0]), np.float16)
        for k in range(0, a.shape[1], 16):
            part = a[:, k:k + 16].astype(np.float64) @ w[:, k:k + 16].T.astype(np.float64)
            r = (r.astype(np.float64) + part).astype(np.float16)
    if relu:
        r = np.maximum(r, np.float16(0))
    return r


def network_forward(snap, x01: np.ndarray, sh16: np.ndarray, accum: str = "fp16_k16") -> np.ndarray:
    """NerfNetwork::inference_mixed_precision_impl -> fp16 [n,4] = (r,g,b raw, density raw)."""
    enc = hash_encode(snap, x01)
    h = _mlp_layer(enc, snap.density_mlp[0], True, accum)
    dens = _mlp_layer(h, snap.density_mlp[1], False, accum)           # [n,16], row 0 = raw density
    rgb_in = np.concatenate([dens, sh16], axis=1)                     # [16 density | 16 SH]
    h = _mlp_layer(rgb_in, snap.rgb_mlp[0], True, accum)
    h = _mlp_layer(h, snap.rgb_mlp[1], True, accum)
    rgb = _mlp_layer(h, snap.rgb_mlp[2], False, accum)
    return np.concatenate([rgb[:, :3], dens[:, :1]], axis=1)


# ----------------------------------------------------------------------------------------------
# full render = Testbed::render(w, h, 1, linear=True)   (src/python_api.cu:123-201)
# ----------------------------------------------------------------------------------------------
SHADE, DEPTH = "Shade", "Depth"


@dataclasses.dataclass
class RenderStats:
    n_rays: int = 0
    n_alive_start: int = 0
    n_samples: int = 0
    n_hit: int = 0


def occupied_box(bitfield: np.ndarray, max_cascade: int, pad: float = 1e-3):
    """Tight NGP-space box around every occupied cell of cascades 0..max_cascade (the cull the CUDA path
    uses: a ray that misses it can never take a sample, so skipping it cannot change any pixel)."""
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    n = NERF_GRID_N_CELLS
    for c in range(max_cascade + 1):
        cells = np.nonzero(np.unpackbits(bitfield[c * (n // 8):(c + 1) * (n // 8)], bitorder="little"))[0].astype(u32)
        if cells.size == 0:
            continue
        xyz = np.stack([morton3D_invert(cells), morton3D_invert(cells >> u32(1)), morton3D_invert(cells >> u32(2))], 1).astype(np.float64)
        size = 2.0 ** c
        lo = np.minimum(lo, 0.5 - size / 2 + size * xyz.min(0) / 128.0)
        hi = np.maximum(hi, 0.5 - size / 2 + size * (xyz.max(0) + 1) / 128.0)
    return (lo - pad).astype(f32), (hi + pad).astype(f32)


def render(snap, bitfield: np.ndarray, vs: ViewSetup, cam_nerf_34, mode: str = SHADE,
           background_color=None, min_transmittance: float = 0.01, accum: str = "fp16_k16",
           plane_dirs: Optional[np.ndarray] = None, pixel_mask: Optional[np.ndarray] = None,
           stats: Optional[RenderStats] = None, both: bool = False, cull_box=None):
    """Returns float32 [H, W, 4] linear premultiplied RGBA, exactly what pyngp hands Python.

    both=True returns (shade, depth) from one march (identical sample sets -- the reference
    marches twice with the same rays, so the two renders see the same samples)."""
    W, H = vs.W, vs.H
    bg = np.asarray(snap.background_color if background_color is None else background_color, f32)
    M = nerf_matrix_to_ngp(cam_nerf_34, snap.dataset_scale, snap.dataset_offset, snap.from_mitsuba)
    R, T = M[:, :3], M[:, 3]
    cam_fwd = R[:, 2].copy()
    if plane_dirs is None:
        plane_dirs = camera_plane_dirs(vs)
    P = W * H
    sel = np.arange(P) if pixel_mask is None else np.nonzero(pixel_mask.reshape(-1))[0]
    dcam = plane_dirs[sel]
    d = (dcam[:, 0:1] * R[:, 0] + dcam[:, 1:2] * R[:, 1] + dcam[:, 2:3] * R[:, 2]).astype(f32)   # mat3 * dir
    o = np.broadcast_to(T, d.shape).astype(f32)
    d = (d / np.sqrt(np.sum(d * d, axis=-1, keepdims=True, dtype=f32))).astype(f32)
    ra_min, ra_max, r2l = snap.render_aabb_min, snap.render_aabb_max, snap.render_aabb_to_local
    ident = np.allclose(r2l, np.eye(3))
    lo_, ld_ = (o, d) if ident else ((o @ r2l.T).astype(f32), (d @ r2l.T).astype(f32))
    t = (np.maximum(ray_aabb(lo_, ld_, ra_min, ra_max), f32(0.0)) + f32(1e-6)).astype(f32)      # testbed_nerf.cu:1468
    p0 = (o + t[:, None] * d).astype(f32)
    lp0 = p0 if ident else (p0 @ r2l.T).astype(f32)
    alive = np.all((lp0 >= ra_min) & (lp0 <= ra_max), axis=-1)
    if cull_box is not None:      # optional, result-preserving: see occupied_box()
        FMAX = np.finfo(f32).max
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            t0_ = ((cull_box[0] - o) / d).astype(f32)
            t1_ = ((cull_box[1] - o) / d).astype(f32)
        tn = np.max(np.minimum(t0_, t1_), axis=-1)
        tf = np.min(np.maximum(t0_, t1_), axis=-1)
        alive &= (tn <= tf) & (tf >= 0) & (tn < FMAX)
    cone = f32(snap.cone_angle_constant)
    with np.errstate(divide="ignore"):
        idir = (f32(1.0) / d).astype(f32)

    n = sel.size
    rgba = np.zeros((n, 4), f32)
    rgba_d = np.zeros((n, 4), f32)
    if stats is not None:
        stats.n_rays += n
    ia = np.nonzero(alive)[0]
    # advance_pos_nerf (testbed_nerf.cu:333-362): Sobol jitter then skip
    seeds = (sel[ia].astype(np.uint64) * np.uint64(786433) & np.uint64(0xFFFFFFFF)).astype(u32)
    tt = advance_n_steps(t[ia], cone, ld_random_val(0, seeds))
    tt = skip_to_occupied(tt, cone, o[ia], d[ia], idir[ia], bitfield, snap.max_cascade, ra_min, ra_max, None if ident else r2l)
    ok = tt < MAX_DEPTH
    ia, tt = ia[ok], tt[ok]
    tcur = np.zeros(n, f32)
    tcur[ia] = tt
    if stats is not None:
        stats.n_alive_start += ia.size
    sh_all = sh_encode_deg4(((d + f32(1.0)) * f32(0.5)).astype(f32) * f32(2.0) - f32(1.0))      # warp then kernel_sh's 2x-1
    aabb_min, aabb_max = snap.aabb_min, snap.aabb_max
    diag = (aabb_max - aabb_min).astype(f32)
    depth_scale = f32(1.0) / f32(snap.dataset_scale)
    step = 0
    while ia.size and step < MARCH_ITER:
        step += 1
        tt = skip_to_occupied(tcur[ia], cone, o[ia], d[ia], idir[ia], bitfield, snap.max_cascade, ra_min, ra_max, None if ident else r2l)
        ok = tt < MAX_DEPTH
        ia, tt = ia[ok], tt[ok]
        if ia.size == 0:
            break
        dt = calc_dt(tt, cone)
        pos = (o[ia] + d[ia] * tt[:, None]).astype(f32)
        wpos = ((pos - aabb_min) / diag).astype(f32)                                   # warp_position
        out = network_forward(snap, wpos, sh_all[ia], accum).astype(f32)
        if stats is not None:
            stats.n_samples += ia.size
        tcur[ia] = (tt + dt).astype(f32)
        # composite_kernel_nerf (testbed_nerf.cu:511-667)
        upos = (aabb_min + wpos * diag).astype(f32)                                    # unwarp_position
        Tr = (f32(1.0) - rgba[ia, 3]).astype(f32)
        dtu = unwarp_dt(warp_dt(dt))
        with np.errstate(over="ignore"):
            dens = np.exp(out[:, 3]).astype(f32)                                       # Exponential density activation
            alpha = (f32(1.0) - np.exp(-(dens * dtu).astype(f32)).astype(f32)).astype(f32)
        wgt = (alpha * Tr).astype(f32)
        col = logistic(out[:, :3])
        rgba[ia, :3] += (col * wgt[:, None]).astype(f32)
        rgba[ia, 3] += wgt
        if mode == DEPTH or both:
            dep = (np.sum(cam_fwd * (upos - o[ia]), axis=-1, dtype=f32) * depth_scale).astype(f32)
            rgba_d[ia, :3] += (dep[:, None] * wgt[:, None]).astype(f32)
            rgba_d[ia, 3] += wgt
        sat = rgba[ia, 3] > f32(f32(1.0) - f32(min_transmittance))
        fin = ia[sat]
        if both or mode == DEPTH:
            rgba_d[fin] = (rgba_d[fin] / rgba_d[fin, 3:4]).astype(f32)
        rgba[fin] = (rgba[fin] / rgba[fin, 3:4]).astype(f32)
        ia = ia[~sat]

    def finish(acc, shade):
        hit = acc[:, 3] > f32(0.001)                                                   # compact_kernel_nerf keep rule
        tmp = np.where(hit[:, None], acc, f32(0.0)).astype(f32)
        if shade:
            tmp[:, :3] = srgb_to_linear(tmp[:, :3])                                    # shade_kernel_nerf (premultiplied!)
        # tonemap_kernel background blend (render_buffer.cu:540-548)
        bgl = srgb_to_linear(bg[:3])
        w = ((f32(1.0) - tmp[:, 3]) * bg[3]).astype(f32)
        tmp[:, :3] = (tmp[:, :3] + bgl[None, :] * w[:, None]).astype(f32)
        tmp[:, 3] = (tmp[:, 3] + w).astype(f32)
        full = np.zeros((P, 4), f32)
        if pixel_mask is not None:   # untouched pixels still get the background blend
            full[:, :3] = bgl * bg[3]
            full[:, 3] = bg[3]
        full[sel] = tmp
        if stats is not None:
            stats.n_hit += int(hit.sum())
        return full.reshape(H, W, 4)

    if both:
        return finish(rgba, True), finish(rgba_d, False)
    return finish(rgba_d, False) if mode == DEPTH else finish(rgba, True)

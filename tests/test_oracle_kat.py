"""Pin the oracle against known-answer vectors host-compiled from the reference's own headers
(generator sources kept beside the JSON as kat_host.cu.txt / kat_cam.cu.txt)."""
import json
import os

import numpy as np
import pytest

from oracle import ngp_oracle as O

f32 = np.float32


@pytest.fixture(scope="module")
def kat(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kat_host.json")))


@pytest.fixture(scope="module")
def kat_cam(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kat_cam.json")))


def test_ld_random_val_bit_exact(kat):
    seeds = np.array([s for s, _ in kat["ld_random_val"]], np.uint32)
    want = np.array([v for _, v in kat["ld_random_val"]], f32)
    got = O.ld_random_val(0, seeds)
    assert np.array_equal(got, want)


def test_pixel_offset(kat):
    for spp, want in enumerate(kat["ld_random_pixel_offset"]):
        np.testing.assert_allclose(O.ld_random_pixel_offset(spp), np.array(want, f32), rtol=0, atol=1e-7)
    assert np.array_equal(O.ld_random_pixel_offset(0), np.array([0.5, 0.5], f32))


def test_sobol_scramble_morton_bit_exact(kat):
    idx = (np.arange(6, dtype=np.uint64) * 2654435761 & 0xFFFFFFFF).astype(np.uint32)
    assert [[int(a), int(b)] for a, b in zip(O.sobol(idx, 0), O.sobol(idx, 1))] == kat["sobol"]
    x = (np.arange(4, dtype=np.uint32) * np.uint32(7919) + np.uint32(1))
    assert [int(v) for v in O.nested_uniform_scramble_base2(x, 0xdeadbeef)] == kat["nested_uniform_scramble_base2"]
    assert [int(O.morton3D(*p)) for p in ((1, 2, 3), (127, 0, 64), (5, 77, 100))] == kat["morton3D"]
    for p in ((1, 2, 3), (127, 0, 64), (5, 77, 100)):
        m = O.morton3D(*p)
        assert (int(O.morton3D_invert(m)), int(O.morton3D_invert(m >> np.uint32(1))), int(O.morton3D_invert(m >> np.uint32(2)))) == p


def test_constants(kat):
    c = kat["constants"]
    assert O.STEPSIZE == f32(c["STEPSIZE"]) and O.MIN_CONE_STEPSIZE == f32(c["MIN_CONE"])
    assert O.MAX_CONE_STEPSIZE == f32(c["MAX_CONE"]) and O.MAX_DEPTH == f32(c["MAX_DEPTH"])


def test_stepping_space(kat):
    for row in kat["stepping"]:
        t, cone = f32(row["t"]), row["cone"]
        np.testing.assert_allclose(O.to_stepping_space(t, cone), f32(row["to"]), rtol=2e-6)
        # dt is a difference of nearly equal numbers: compare against an ulp of t
        assert abs(float(O.calc_dt(t, cone)) - row["dt"]) <= 4 * np.spacing(f32(max(t, 1e-3)))
        np.testing.assert_allclose(O.advance_n_steps(t, cone, f32(2.5)), f32(row["adv2.5"]), rtol=1e-6)


def test_warp_dt(kat):
    np.testing.assert_allclose(O.warp_dt(f32([0.002, 0.1])), kat["warp_dt"], rtol=1e-6)
    np.testing.assert_allclose([O.unwarp_dt(f32(0.25)), O.unwarp_dt(O.warp_dt(f32(0.0123)))], kat["unwarp_dt"], rtol=1e-6)


PTS = np.array([[.5, .5, .5], [.9, .5, .5], [1.2, .1, .5], [-.4, .5, 1.4], [.5, 2.6, .5]], f32)


def test_mip_and_cascade_index_bit_exact(kat):
    assert [int(m) for m in O.mip_from_pos(PTS)] == kat["mip_from_pos"]
    got = [[int(O.cascaded_grid_idx_at(p[None], np.array([m]))[0]) for m in (0, 1)] for p in PTS]
    assert got == kat["cascaded_grid_idx_at"]


def test_dda(kat):
    p = np.array([[0.31, 0.52, 0.77]], f32)
    d = np.array([[0.3, -0.8, 0.52]], f32)
    d = (d / np.sqrt((d * d).sum(dtype=f32))).astype(f32)
    idir = (f32(1) / d).astype(f32)
    for mip in range(3):
        res = np.ldexp(f32(128), -mip)
        np.testing.assert_allclose(O.distance_to_next_voxel(p, d, idir, np.array([res], f32))[0], kat["distance_to_next_voxel"][mip], rtol=2e-5)
        got = O.advance_to_next_voxel(np.array([0.4], f32), 1 / 256, p, d, idir, np.array([mip]))[0]
        np.testing.assert_allclose(got, kat["advance_to_next_voxel"][mip], rtol=1e-6)


def test_srgb_and_activations(kat):
    np.testing.assert_allclose([O.linear_to_srgb_cpp(f32(0.002)), O.linear_to_srgb_cpp(f32(0.5)),
                                O.srgb_to_linear(f32(0.03)), O.srgb_to_linear(f32(0.5))], kat["srgb"], rtol=1e-6)
    np.testing.assert_allclose(O.logistic(f32(0.3)), kat["network_to"]["logistic(0.3)"], rtol=1e-6)


def test_lens_undistortion(kat):
    lens = np.array([0.096692, -0.166479, -0.000194, 0.002049], f32)
    u, v = O.iterative_opencv_lens_undistortion(lens, np.array([0.41], f32), np.array([-0.27], f32))
    np.testing.assert_allclose([u[0], v[0]], kat["opencv_undistort"]["out"], rtol=0, atol=2e-7)


def test_ray_aabb(kat):
    d = np.array([[0.1, 0.05, -1.0]], f32)
    d = (d / np.sqrt((d * d).sum(dtype=f32))).astype(f32)
    t = O.ray_aabb(np.array([[0.5, 0.3, 3.0]], f32), d, f32(-0.5) * np.ones(3, f32), f32(1.5) * np.ones(3, f32))
    np.testing.assert_allclose(t[0], kat["ray_intersect"][0], rtol=1e-6)


class _Snap:
    """shopping-scene dataset block (reference configs/shopping_demo.json:48-67)."""
    from dream2real_b200.ingp import ViewMeta
    views = [ViewMeta(np.array([924.66912, 926.49735], f32), np.array([654.51953 / 1280, 355.18523 / 720], f32),
                      np.array([1280, 720], np.int32), "OpenCV", np.array([0.096692, -0.166479, -0.000194, 0.002049], f32))]
    fov_axis = 1
    zoom = 1.0
    dataset_scale = 1.0
    dataset_offset = np.array([0.0, 0.3, 0.5], f32)
    from_mitsuba = False


def test_camera_pipeline(kat_cam):
    M = O.nerf_matrix_to_ngp(np.array(kat_cam["nerf_cam"], f32), 1.0, [0.0, 0.3, 0.5])
    np.testing.assert_allclose(M, np.array(kat_cam["ngp_cam"], f32), rtol=0, atol=1e-7)
    for res in (336, 800):
        ref = kat_cam[f"res{res}"]
        vs = O.view_setup(_Snap, 0, res, res)
        np.testing.assert_allclose(vs.focal, ref["focal"], rtol=1e-6)
        np.testing.assert_allclose(vs.screen_center, ref["screen_center"], rtol=1e-6)
        dirs = O.camera_plane_dirs(vs)
        for ray in ref["rays"]:
            x, y = ray["px"]
            dc = dirs[x + res * y]
            d = dc[0] * M[:, 0] + dc[1] * M[:, 1] + dc[2] * M[:, 2]
            d = d / np.linalg.norm(d)
            np.testing.assert_allclose(d, ray["d"], rtol=0, atol=3e-7)

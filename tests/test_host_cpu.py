"""CPU tests (no GPU): host-side mirrors of the reference interface vs the oracle, the C-ABI library's
exported symbols, snapshot I/O round trips, pose sharding with a world_size-2 gloo group."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/d2r_b200.h <-> libd2r_b200.so <-> the ctypes table must agree (no compute calls here)."""
    from dream2real_b200 import _native as N
    hdr = open(os.path.join(ROOT, "include", "d2r_b200.h")).read()
    declared = set(re.findall(r"\b(d2r_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    assert os.path.exists(N.LIB_PATH), "build the library first: python dream2real_b200/csrc/build.py"
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    lib.d2r_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.d2r_version()


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dream2real_b200 import testbed
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        testbed.Testbed(testbed.TestbedMode.Nerf)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dream2real_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_snapshot_roundtrip(tmp_path):
    from dream2real_b200 import ingp, synth
    sc = synth.make_scene("pool_triangle", str(tmp_path), log2_hashmap_size=12, seed=3)
    s = ingp.load_snapshot(os.path.join(sc["dir"], "fg_base.ingp"))
    assert s.aabb_scale == 2 and s.max_cascade == 1 and s.cone_angle_constant == 1 / 256
    assert s.grid.log2_hashmap_size == 12 and s.params.size == 10240 + s.grid.n_params
    assert s.dataset_scale == 1.0 and np.allclose(s.dataset_offset, [0, 0.3, 0.5])
    assert s.views[0].lens_mode == "OpenCV" and np.allclose(s.views[0].lens_params, [0.096692, -0.166479, -0.000194, 0.002049])
    assert np.allclose(s.background_color, 0)
    p2 = str(tmp_path / "again.ingp")
    ingp.save_snapshot(p2, s.config)
    s2 = ingp.load_snapshot(p2)
    assert np.array_equal(s.params, s2.params) and np.array_equal(s.density_grid, s2.density_grid)


def test_snapshot_errors(tmp_path):
    from dream2real_b200 import ingp, synth
    sc = synth.make_scene("shopping", str(tmp_path), log2_hashmap_size=12)
    s = ingp.load_snapshot(os.path.join(sc["dir"], "fg_base.ingp"))
    cfg = dict(s.config)
    bad = dict(cfg, snapshot=dict(cfg["snapshot"], version=0))
    with pytest.raises(RuntimeError, match="old format"):
        ingp.decode_snapshot(bad)
    bad = dict(cfg, snapshot=dict(cfg["snapshot"], density_grid_size=64))
    with pytest.raises(RuntimeError, match="Incompatible grid size"):
        ingp.decode_snapshot(bad)
    bad = {k: v for k, v in cfg.items() if k != "snapshot"}
    with pytest.raises(RuntimeError, match="does not contain a snapshot"):
        ingp.decode_snapshot(bad)
    bad = dict(cfg, snapshot=dict(cfg["snapshot"], density_grid_binary=b"\x00" * 10))
    with pytest.raises(RuntimeError, match="cascades"):
        ingp.decode_snapshot(bad)


class _TM:
    def __init__(self, centre):
        import types
        self.scene_model = types.SimpleNamespace(scene_centre=torch.tensor(centre), device=torch.device("cpu"))


@pytest.mark.parametrize("scene_type,res", [(0, [5, 4, 2, 1, 1, 1]), (3, [3, 3, 1, 1, 1, 1]), (1, [2, 2, 2, 3, 2, 2])])
def test_pose_grid_vs_oracle(scene_type, res):
    from dream2real_b200.vision_3d.obj_pose_opt import sample_poses_grid
    from oracle import post_oracle as PO
    c = [0.5, 0.0, 0.035]
    got = sample_poses_grid(_TM(c), res, scene_type=scene_type)
    ref = PO.sample_poses_grid(c, res, scene_type)
    assert got.shape == (int(np.prod(res)), 16) and torch.equal(got, ref)
    # x slowest, z-rotation fastest; rotations are orthonormal
    R = got.view(-1, 4, 4)[:, :3, :3]
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand_as(R), atol=1e-6)
    assert torch.all(got.view(-1, 4, 4)[1:, 0, 3] >= got.view(-1, 4, 4)[:-1, 0, 3])
    with pytest.raises(NotImplementedError):
        sample_poses_grid(_TM(c), res, scene_type=2)


def test_euler_xyz_known_answer():
    from dream2real_b200.vision_3d.obj_pose_opt import euler_angles_to_matrix_xyz
    from scipy.spatial.transform import Rotation
    e = torch.tensor([[0.3, -1.1, 2.0], [-3.0, 0.4, 1.5]], dtype=torch.float64)
    ref = Rotation.from_euler("XYZ", e.numpy()).as_matrix()     # intrinsic XYZ == Rx @ Ry @ Rz
    assert np.allclose(euler_angles_to_matrix_xyz(e).numpy(), ref, atol=1e-12)


def test_frame_conversions():
    from dream2real_b200.reconstruction.combined_rendering import convert_virtual_pose
    from dream2real_b200.utils.accio2ngp import converter
    from oracle import post_oracle as PO
    rng = np.random.default_rng(0)
    T = np.tile(np.eye(4), (3, 1, 1))
    T[:, :3, :] = rng.normal(size=(3, 3, 4))
    assert np.array_equal(converter(T), PO.converter(T)) and not np.shares_memory(converter(T), T)
    assert np.array_equal(converter(converter(T)), T)
    a, b, c = T
    assert np.allclose(convert_virtual_pose(a, b, c), PO.convert_virtual_pose(a, b, c))
    assert np.allclose(convert_virtual_pose(a, a, c), c)                    # unmoved object -> real camera
    assert np.allclose(convert_virtual_pose(a, b, c), a @ np.linalg.inv(b) @ c)


@pytest.mark.parametrize("res", [[6, 5, 1, 1, 1, 1], [4, 4, 2, 1, 1, 3]])
def test_smoothing_vs_torchvision_oracle(res):
    from dream2real_b200.vision_3d.geometry_utils import spatially_smooth_heatmap
    from oracle import post_oracle as PO
    torch.manual_seed(1)
    s = torch.rand(int(np.prod(res))) + 0.5
    s[torch.rand_like(s) < 0.2] = 0                       # physics-invalid poses
    got, ref = spatially_smooth_heatmap(s, res), PO.spatially_smooth_heatmap(s, res)
    assert torch.allclose(got, ref, atol=1e-6) and torch.equal(got == 0, s == 0)


def test_preprocess_oracle_vs_hf_pil_processor():
    """oracle clip_preprocess == transformers' PIL-backed CLIP image processor (the 4.27.3 behaviour)."""
    from oracle import post_oracle as PO
    try:
        from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil as Proc
    except Exception:
        pytest.skip("PIL-backed CLIP image processor not available in this transformers build")
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(2, 90, 90, 3), dtype=np.uint8)
    proc = Proc(size={"shortest_edge": 64}, crop_size={"height": 64, "width": 64})
    ref = proc(images=[i for i in imgs], return_tensors="np")["pixel_values"]
    got = PO.clip_preprocess(imgs, 64)
    assert np.abs(np.asarray(ref) - got).max() < 1e-6


def test_score_normalisation_oracle():
    from oracle import post_oracle as PO
    L = torch.tensor([[2.0, 4.0, 8.0], [3.0, 1.0, 2.0]])
    assert torch.allclose(PO.normalise_scores(L, 1), torch.tensor([2.0 / 6.0, 3.0 / 1.5]))
    assert torch.allclose(PO.normalise_scores(L, 3), L.mean(1))
    assert torch.allclose(PO.normalise_scores(L, 2), torch.tensor([3.0 / 8.0, 2.0 / 2.0]))


def test_shard_bounds():
    from dream2real_b200.clip_scoring import shard_bounds
    for n in (0, 1, 7, 8, 9, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == (n + world - 1) // world if n else True


def _gloo_worker(rank, world, port, n, q):
    import torch.distributed as dist
    from dream2real_b200.clip_scoring import gather_scores, shard_bounds, shard_indices
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(n, dtype=torch.float32) * 0.5 + 1
    lo, hi = shard_bounds(n, world, rank)
    ok = torch.equal(gather_scores(full[lo:hi].clone(), n, world, rank), full)
    for mode in ("contiguous", "strided"):      # both splits come back in the candidates' original order
        idx = shard_indices(n, world, rank, mode)
        ok = ok and torch.equal(gather_scores(full[idx].clone(), n, world, rank, mode=mode), full)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 9, 64])
def test_score_all_gather_world2_gloo(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert sorted(res) == [(0, True), (1, True)]


def test_composite_oracle_semantics():
    from oracle import post_oracle as PO
    bg = np.zeros((2, 2, 4), np.float32)
    bg[..., :3] = 0.2
    bg[..., 3] = 1.0
    fg = np.zeros((2, 2, 4), np.float32)
    fg[0, 0] = [0.5, 0.25, 0.0, 1.0]        # opaque, in front
    fg[0, 1] = [0.5, 0.25, 0.0, 1.0]        # opaque, behind the background
    fg[1, 0] = [0.2, 0.1, 0.0, 0.4]         # in front but alpha*255 < 130 -> black
    fg_d = np.array([[0.5, 3.0], [0.5, 0.0]], np.float32)
    bg_d = np.array([[1.0, 1.0], [1.0, 0.01]], np.float32)   # 0.01 < 0.05 -> 100
    out = PO.composite(bg, bg_d, fg, fg_d)
    s = lambda x: int(np.clip(1.055 * x ** (1 / 2.4) - 0.055, 0, 1) * 255 + 0.5)
    assert tuple(out[0, 0]) == (s(0.5), s(0.25), 0)
    assert tuple(out[0, 1]) == (s(0.2),) * 3
    assert tuple(out[1, 0]) == (0, 0, 0)
    assert tuple(out[1, 1]) == (s(0.2),) * 3                # fg depth 0 -> 100, not < bg 100


def test_optimise_pose_grid_argument_errors():
    """Argument checks that must fire before any device work (reference: use_vis_pcds selects the point-cloud ablation
    renderer, clip_scoring.py:118-131, which is not on the accelerated path)."""
    import pytest
    from dream2real_b200 import clip_scoring
    with pytest.raises(NotImplementedError):
        clip_scoring.optimise_pose_grid(None, None, [0], None, "/tmp", use_vis_pcds=True)
    with pytest.raises(ValueError):
        clip_scoring.optimise_pose_grid(None, None, [0, 1], None, "/tmp", multi_view="median")


def test_background_depth_rectification_matches_reference_recipe():
    """renderer.render_background's host recipe (centre crop, cv2 INTER_CUBIC resize, 100 m where the mask is 0;
    combined_rendering.py:104-111,166-209) against the literal 4-channel formulation of the reference."""
    import cv2
    import torch
    from dream2real_b200.reconstruction.combined_rendering import renderer
    r = renderer.__new__(renderer)
    r.resolution = [96, 96]
    g = torch.Generator().manual_seed(3)
    depth = (torch.rand(72, 128, generator=g) * 2).half()
    mask = torch.rand(72, 128, generator=g) > 0.2
    # reference formulation
    crop = depth.numpy()[:, 28:100].astype(np.float32)
    d4 = np.repeat(np.expand_dims(cv2.resize(crop, (96, 96), interpolation=cv2.INTER_CUBIC), axis=2), 4, axis=2)
    m = cv2.resize(mask.numpy()[:, 28:100].astype(np.uint8), (96, 96), interpolation=cv2.INTER_CUBIC)
    d4[m == 0, 0] = 100
    # ours
    d1 = r._rectify_depth_1ch(depth, r.resolution)
    np.putmask(d1, r.rectify_mask(mask, r.resolution) == 0, np.float32(100))
    assert np.array_equal(d1, d4[..., 0])
    assert np.array_equal(r.rectify_depth(depth, r.resolution)[..., 2], cv2.resize(crop, (96, 96), interpolation=cv2.INTER_CUBIC))


def test_cpu_arm_banded_pipeline_matches_per_candidate_pipeline():
    """bench.py's CPU arms: `--impl reference` deals every candidate's pixel rows out over the worker processes (so that a step
    can be shorter than one candidate on one core), the `cpu_baseline` leg renders one candidate per worker.  Same oracle, same
    rays, same composite -> identical scores."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    grid = [4, 4, 1, 1, 1, 1]
    whole = bench.cpu_pipeline("shopping", grid, 16, 96, "ViT-B/32", 12, 2, split=1)
    a = np.asarray(whole([1, 10]))
    whole.close()
    banded = bench.cpu_pipeline("shopping", grid, 16, 96, "ViT-B/32", 12, 2, split=3)
    b = np.asarray(banded([1, 10]))
    banded.close()
    assert a.shape == (2,) and np.isfinite(a).all()
    assert np.array_equal(a, b)

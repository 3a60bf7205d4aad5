"""-m gpu: the pre-render physics filter (reference vision_3d/physics_utils.py:232-378) -- the one-launch
CUDA kernel behind create_unsupcol_check against oracle/phys_oracle.py on the three scene families, bit for bit (a bool mask),
plus the properties the definition implies on scenes whose geometry is known analytically."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _world(name, d, log2=12):
    import torch
    from dream2real_b200 import synth
    scene = synth.make_scene(name, d, log2_hashmap_size=log2, seed=11)
    return scene, synth.SyntheticTaskModel(scene, "g", None, torch.device("cuda"))


@pytest.mark.parametrize("name,sample_res,embodied", [
    ("shopping", [9, 9, 8, 1, 1, 1], False),
    ("pool_triangle", [10, 10, 6, 1, 1, 1], False),
    ("shelf", [4, 2, 6, 2, 2, 2], True),          # 6-DoF: orientation-uniqueness and regrasp masks come into play
])
def test_unsupcol_check_matches_oracle(tmp_path, name, sample_res, embodied):
    import torch
    from dream2real_b200.vision_3d.obj_pose_opt import sample_poses_grid
    from dream2real_b200.vision_3d.physics_utils import create_unsupcol_check, occupied_points_world
    from oracle import ngp_oracle as O
    from oracle import phys_oracle as PH
    scene, tm = _world(name, str(tmp_path))
    poses = sample_poses_grid(tm, sample_res, scene_type=scene["scene_type"])
    check, statics, movables = create_unsupcol_check(None, tm, sample_res, embodied)
    assert statics == [] and movables == []
    valid0 = torch.ones(poses.shape[0], dtype=torch.bool, device=poses.device)
    got = check(poses, tm, valid0).cpu().numpy()
    assert got.dtype == bool and got.shape == (poses.shape[0],)
    # the oracle, from the snapshots
    fg, bg = tm.movable_obj.vis_model.snapshot, tm.task_bground_obj.vis_model.snapshot
    fgb, _ = O.build_bitfield(fg.density_grid, fg.max_cascade)
    bgb, _ = O.build_bitfield(bg.density_grid, bg.max_cascade)
    pts = PH.ngp_to_world(PH.occupied_points_ngp(fgb, fg.max_cascade), fg.dataset_scale, fg.dataset_offset)
    assert np.array_equal(pts, occupied_points_world(tm.movable_obj.vis_model))
    ref = PH.unsupcol_check(poses.cpu().numpy(), tm.movable_obj.pose.cpu().numpy(), pts, bgb, bg.max_cascade, bg.dataset_scale, bg.dataset_offset,
                            float(tm.scene_model.scene_centre[2]), sample_res, np.ones(poses.shape[0], bool), disallow_regrasp=embodied)
    print(f"{name}: {poses.shape[0]} poses, {int(got.sum())} valid (oracle {int(ref.sum())}), object proxy {pts.shape[0]} points")
    assert np.array_equal(got, ref)
    assert 0 < got.sum() < got.size          # the grid spans colliding, supported and floating poses


def test_physics_definition_on_known_geometry(tmp_path):
    """Shopping stand-in: a 4 cm sphere over a table slab (z in [-0.03, 0]) with boxes on it.  Resting just above the table is valid;
    the same spot 10 cm up is unsupported; inside a box is in collision; validity never depends on poses elsewhere in the batch."""
    import torch
    from dream2real_b200.vision_3d.physics_utils import create_unsupcol_check
    scene, tm = _world("shopping", str(tmp_path))
    check, _, _ = create_unsupcol_check(None, tm, [1, 1, 1, 1, 1, 1], False)

    def pose(x, y, z):
        p = torch.eye(4)
        p[:3, 3] = torch.tensor([x, y, z])
        return p.reshape(1, 16)
    free_xy = (0.50, -0.02)                     # table top, away from the boxes (synth.SCENES["shopping"])
    batch = torch.cat([pose(*free_xy, 0.055), pose(*free_xy, 0.16), pose(0.36, -0.16, 0.05), pose(1.2, 0.0, 0.02)]).to(tm.scene_model.device)
    v = check(batch, tm, torch.ones(4, dtype=torch.bool, device=batch.device)).cpu().tolist()
    print("resting, floating, inside a box, beside the table below its plane:", v)
    assert v[0] is True and v[1] is False and v[2] is False
    assert v[3] is True        # free of collisions and below scene_centre z (0.035): the reference counts it as supported (:333-335)
    # independent of the batch, and valid_so_far is respected
    for i in range(4):
        assert check(batch[i:i + 1], tm, torch.ones(1, dtype=torch.bool, device=batch.device)).cpu().tolist() == [v[i]]
    assert check(batch, tm, torch.zeros(4, dtype=torch.bool, device=batch.device)).sum().item() == 0


def test_physics_filter_feeds_optimise_pose_grid(tmp_path):
    """the mask is what optimise_pose_grid consumes as phys_check (clip_scoring.py:109-112): invalid poses score 0"""
    import torch
    from dream2real_b200 import clip_scoring
    from dream2real_b200.reconstruction.combined_rendering import renderer
    from dream2real_b200.vision_3d.physics_utils import create_unsupcol_check
    from test_e2e_gpu import _tiny_clip
    d = str(tmp_path)
    scene, tm = _world("shopping", d)
    tm.goal_caption, tm.norm_captions = "goal", ["norm"]
    sample_res = [9, 9, 8, 1, 1, 1]
    check, _, _ = create_unsupcol_check(None, tm, sample_res, False)
    model = _tiny_clip(3)
    ids = torch.randint(3, 900, (2, 6))
    ids[:, -1] = 2
    r = renderer(d, tm, resolution=64)
    best, poses, scores = clip_scoring.optimise_pose_grid(r, tm.depths[:1], [0], tm, d, sample_res=sample_res, phys_check=check, scene_type=3,
                                                          smoothing=False, clip_model=model, text_inputs={"input_ids": ids}, save_renders=False)
    mask = check(poses, tm, torch.ones(poses.shape[0], dtype=torch.bool, device=poses.device)).cpu()
    assert torch.equal(scores != 0, mask) and 0 < int(mask.sum()) < mask.numel()
    assert bool(mask[int(torch.argmax(scores))])

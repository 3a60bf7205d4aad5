"""Pin the oracle against the REAL reference renderer: pyngp (built from /root/reference, run on a B200 by
tests/golden/make_golden_pyngp.py).  fox_a2_small_packed.ingp is the pyngp-trained snapshot the goldens were
rendered from, re-packed by tests/golden/repack_fixture.py (bitfield-preserving)."""
import os

import numpy as np
import pytest

from dream2real_b200 import ingp
from oracle import ngp_oracle as O


@pytest.fixture(scope="module")
def fox(golden_dir):
    snap = ingp.load_snapshot(os.path.join(golden_dir, "fox_a2_small_packed.ingp"))
    bits, thr = O.build_bitfield(snap.density_grid, snap.max_cascade)
    return snap, bits


def test_snapshot_metadata_matches_pyngp(fox, golden_dir):
    import json
    snap, _ = fox
    info = json.load(open(os.path.join(golden_dir, "fox_a2_small_info.json")))
    assert snap.params.size == info["n_params"]
    assert snap.grid.n_params == info["n_encoding_params"]
    assert np.allclose(snap.aabb_min, info["aabb"][0]) and np.allclose(snap.aabb_max, info["aabb"][1])
    assert np.allclose(snap.render_aabb_min, info["render_aabb"][0])
    assert snap.cone_angle_constant == pytest.approx(info["cone_angle_constant"])
    assert snap.dataset_scale == pytest.approx(info["dataset_scale"])
    assert snap.aabb_scale == info["dataset_aabb_scale"] and len(snap.views) == info["n_images"]
    assert abs(snap.grid.per_level_scale - 2.20818) < 1e-5          # pyngp log line "b=2.20818"


@pytest.mark.parametrize("cam", [0, 4])
def test_oracle_matches_pyngp_48(fox, golden_dir, cam):
    snap, bits = fox
    g = np.load(os.path.join(golden_dir, "fox_a2_small_48_bg0.npz"))
    vs = O.view_setup(snap, 0, 48, 48)
    st = O.RenderStats()
    shade, depth = O.render(snap, bits, vs, g["cams"][cam][:3], both=True, background_color=[0, 0, 0, 0], stats=st)
    e = np.abs(shade - g["Shade"][cam])
    # reference = fp16 tensor cores + --use_fast_math; tolerance = north-star 1e-3 on 99.9 % of values
    assert np.percentile(e, 99.9) < 1e-3 and e.mean() < 1e-4 and e.max() < 2e-2, (e.max(), e.mean())
    ed = np.abs(depth[..., 0] - g["Depth"][cam][..., 0])
    assert np.percentile(ed, 99.9) < 5e-3 * max(1.0, float(g["Depth"][cam].max()))
    # alpha channel of the Depth render equals that of the Shade render (same samples)
    assert np.abs(depth[..., 3] - g["Depth"][cam][..., 3]).max() < 2e-2
    # the reference's own step counter (Cost mode = n_steps/128) agrees with our sample count to 1 %
    steps_ref = float(g["Cost"][cam][..., 0].sum() * 128)
    assert abs(st.n_samples - steps_ref) / steps_ref < 0.01


def test_opaque_background_blend(fox, golden_dir):
    snap, bits = fox
    g = np.load(os.path.join(golden_dir, "fox_a2_small_96_bg1.npz"))
    vs = O.view_setup(snap, 0, 96, 96)
    shade = O.render(snap, bits, vs, g["cams"][0][:3], mode=O.SHADE, background_color=[0, 0, 0, 1])
    assert np.allclose(shade[..., 3], 1.0, atol=1e-6) and np.allclose(g["Shade"][0][..., 3], 1.0, atol=1e-6)
    e = np.abs(shade - g["Shade"][0])
    assert np.percentile(e, 99.9) < 1e-3 and e.mean() < 1e-4

"""Visual-model loader (reference reconstruction/ngp_visual_model.py:20-59)."""
import os

from .. import pyngp as ngp


def get_vis_ngps(rgbs, movable_masks, scene_type, use_cache=False, data_dir=None, fg=True, render_distract=False):
    """use_cache=True: load {data_dir}/fg_base.ingp (fg) or bg_base.ingp and return the Testbed handle.
    The training branch of the reference (writes RGBA masks, fine-tunes a NeRF) is upstream of the
    scoring path and out of scope here."""
    if not use_cache:
        raise NotImplementedError("NeRF (re)training is not part of the imagination-and-scoring path: "
                                  "train with the reference and pass use_cache=True")
    print('Using cached fg model for movable object')
    testbed = ngp.Testbed(ngp.TestbedMode.Nerf)
    testbed.load_snapshot(os.path.join(data_dir, 'fg_base.ingp' if fg else 'bg_base.ingp'))
    return testbed


get_visual_model = get_vis_ngps   # north-star alias

"""Pose-grid scoring -- the B200 replacement for reference clip_scoring.py:71-235.

`optimise_pose_grid` keeps the reference signature (extra keyword-only knobs have defaults) and
return value.  Underneath: batched fused render+composite (CUDA uint8 tensor, no host round trip) ->
rot90 + PIL-exact preprocessing kernel -> tcgen05 ViT forward -> score kernel; candidate poses are
sharded contiguously over the ranks of an initialised torch.distributed NCCL group and the score
vector is all-gathered once.
"""
import os

import numpy as np
import torch

from . import clip as d2r_clip
from .utils import accio2ngp
from .vision_3d.geometry_utils import spatially_smooth_heatmap
from .vision_3d.obj_pose_opt import sample_poses_grid
from .vision_3d.virtual_cam_pose_sample import get_virtual_cam_poses

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
CLIP_RES = 336
CLIP_NAME = "openai/clip-vit-large-patch14-336"      # clip_scoring.py:150-151

_clip_cache = {}


def _load_clip(clip_model, clip_processor, need_processor=True):
    """clip_scoring.py:150-151; D2R_CLIP_PATH points at a local copy of the checkpoint (there is no network here)."""
    path = os.environ.get("D2R_CLIP_PATH", CLIP_NAME)
    if clip_model is None:
        from transformers import CLIPModel
        clip_model = CLIPModel.from_pretrained(path).eval()
    if clip_processor is None and need_processor:      # also when the caller brought a model but neither a processor nor text_inputs
        from transformers import CLIPProcessor
        clip_processor = CLIPProcessor.from_pretrained(path)
    return clip_model, clip_processor


def _vision_for(clip_model, device, max_batch):
    """One vision tower on the device per HF model; the entry keeps the model alive, so its id() cannot be recycled by another."""
    key = (id(clip_model), int(device), int(max_batch))
    hit = _clip_cache.get(key)
    if hit is None or hit[0] is not clip_model:
        _clip_cache.clear()
        hit = (clip_model, d2r_clip.ClipVision(clip_model, max_batch=max_batch, device=device))
        _clip_cache[key] = hit
    return hit[1]


def shard_bounds(n, world_size, rank):
    """Contiguous shard [lo, hi) of n candidates for `rank`: ceil(n / R) per rank, last shards may be short/empty."""
    per = (n + world_size - 1) // world_size
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard_indices(n, world_size, rank, mode="strided"):
    """Indices of the candidates rank `rank` scores.  "contiguous": the block [lo, hi) of shard_bounds;
    "strided" (default): rank, rank + R, rank + 2R, ... -- a pose grid is ordered x slowest, so contiguous blocks are x-slabs
    whose objects sit at different distances from the camera and cost different amounts (the shelf grid sharded 8 ways: the
    slowest slab took 1.6x the mean); interleaving gives every rank the same mix."""
    if mode == "contiguous":
        lo, hi = shard_bounds(n, world_size, rank)
        return torch.arange(lo, hi)
    if mode != "strided":
        raise ValueError("shard mode must be 'strided' or 'contiguous'")
    return torch.arange(rank, n, world_size)


def gather_scores(local_scores, n_total, world_size, rank, group=None, mode="contiguous"):
    """ONE all-gather of the per-rank score shard (padded to ceil(n/R)) -> full [n_total] vector on every rank, in the
    candidates' original order for either sharding mode."""
    import torch.distributed as dist
    per = (n_total + world_size - 1) // world_size
    buf = torch.zeros(per, dtype=torch.float32, device=local_scores.device)
    buf[: local_scores.numel()] = local_scores
    out = torch.empty(per * world_size, dtype=torch.float32, device=local_scores.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    if mode == "strided":      # out[r, j] is candidate j * R + r
        out = out.view(world_size, per).t().reshape(-1)
    return out[:n_total]


def score_renders(renders_u8, clip_vision, txt_embeds, n_goal=1, bg_u8=None, rects=None):
    """uint8 CUDA [K,H,W,3] (not yet rotated) -> ratio scores [K] (clip_scoring.py:145-203).  bg_u8 / rects: what
    renderer.render recorded about its frames (renderer.last_bg_u8 / last_rects), lets preprocessing skip the pixels
    every frame shares with the background."""
    emb = clip_vision.encode_images(renders_u8, rot90=True, bg_u8=bg_u8, rects=rects)
    return clip_vision.score(emb, txt_embeds, n_goal=n_goal)


def optimise_pose_grid(renderer,
                       depths_gt,
                       render_cam_pose_idx,
                       task_model,
                       data_dir,
                       sample_res=None,
                       phys_check=None,
                       use_templates=False,
                       scene_type=0,
                       use_vis_pcds=False,
                       use_cache_renders=False,
                       smoothing=True,
                       physics_only=False,
                       *,
                       clip_model=None,
                       clip_processor=None,
                       text_inputs=None,
                       save_renders=True,
                       show_best=False,
                       clip_batch_size=512,
                       multi_view="mean",
                       shard="strided"):
    """Reference signature and return value (clip_scoring.py:71-235).  Keyword-only extras: clip_model / clip_processor /
    text_inputs (inject a loaded CLIP instead of downloading one), save_renders, show_best, clip_batch_size, and
    multi_view: with more than one entry in render_cam_pose_idx the reference indexes K*L renders as if they were K
    (clip_scoring.py:205-206 -- it only ever runs with one view); here the L per-view scores of a pose are averaged
    ("mean") or the best view is taken ("max"); shard: how the valid poses are split over the ranks of an initialised
    torch.distributed group ("strided" balances the ranks, "contiguous" is the block split; shard_indices)."""
    if use_vis_pcds:
        raise NotImplementedError("the point-cloud ablation renderer is not on the accelerated path")
    if multi_view not in ("mean", "max"):
        raise ValueError("multi_view must be 'mean' or 'max'")
    n_views = 1 if use_cache_renders else len(render_cam_pose_idx)
    if sample_res is None:
        sample_res = [40, 40, 1, 1, 1, 1]
    device = torch.device("cuda", torch.cuda.current_device())
    pose_batch = sample_poses_grid(task_model, sample_res, scene_type=scene_type)
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    world = torch.distributed.get_world_size() if dist_on else 1
    rank = torch.distributed.get_rank() if dist_on else 0

    renders = None
    stream_args = None
    if use_cache_renders:
        import cv2
        print('Using cached renders')
        old_pose_scores = torch.from_numpy(np.loadtxt(os.path.join(data_dir, 'pose_scores.txt')))
        valid_idxs = torch.nonzero(old_pose_scores).squeeze(-1)
        valid_poses = pose_batch[valid_idxs]
        render_dir = os.path.join(data_dir, 'cb_render')
        imgs = [cv2.cvtColor(cv2.imread(os.path.join(render_dir, f)), cv2.COLOR_BGR2RGB) for f in sorted(os.listdir(render_dir))]
        assert len(imgs) == valid_poses.shape[0], f'Expected {valid_poses.shape[0]} renders, got {len(imgs)}. Try running without use_cache_renders.'
        renders = torch.from_numpy(np.stack(imgs)).to(device)
        lo, hi = 0, len(imgs)
        world = 1
    else:
        print('Using CLIP templates' if use_templates else 'Not using CLIP templates')
        print('Running pre-render checks...')
        valid_so_far = torch.ones(pose_batch.shape[0]).bool().to(pose_batch.device)
        is_valid = phys_check(pose_batch, task_model, valid_so_far)
        valid_idxs = torch.nonzero(is_valid).squeeze(-1)
        valid_poses = pose_batch[valid_idxs]
        print(f'Of {pose_batch.shape[0]} sampled poses, {valid_idxs.shape[0]} passed pre-render checks ({100 * valid_idxs.shape[0] / pose_batch.shape[0]:.2f}%).')
        if valid_idxs.shape[0] == 0:
            print('No poses passed pre-render checks. Exiting.')
            raise Exception
        if physics_only:
            print('Physics only method')
            pick = torch.randint(valid_idxs.shape[0], (1,))
            if dist_on:      # every rank returns the same pose: rank 0's draw
                pick = pick.to(device)
                torch.distributed.broadcast(pick, src=0)
            best_pose_idx = int(pick.item())
            return valid_poses[best_pose_idx].view(4, 4), pose_batch, torch.ones(pose_batch.shape[0])
        render_poses = get_virtual_cam_poses(task_model, render_cam_pose_idx)
        print('Rendering images from ngp...')
        render_poses_ngp = accio2ngp.converter(render_poses)
        valid_poses_ngp = accio2ngp.converter(valid_poses.cpu().numpy().reshape(-1, 4, 4))
        mine = shard_indices(valid_poses_ngp.shape[0], world, rank, shard).numpy()
        stream_args = (valid_poses_ngp, render_poses_ngp)

    print('Evaluating rendered images using CLIP...')
    clip_model, clip_processor = _load_clip(clip_model, clip_processor, need_processor=text_inputs is None)
    goal_caption = task_model.goal_caption
    norm_captions = task_model.norm_captions
    n_goal = 1
    if use_templates:
        from .clip_text_templates import CLIP_TEMPLATES
        captions = [t.format(goal_caption) for t in CLIP_TEMPLATES]
        n_goal = len(CLIP_TEMPLATES)
        if norm_captions is not None:
            for c in norm_captions:
                captions += [t.format(c) for t in CLIP_TEMPLATES]
    else:
        captions = [goal_caption] if norm_captions is None else [goal_caption] + list(norm_captions)
    if text_inputs is None:
        text_inputs = clip_processor(text=captions, return_tensors="pt", padding=True)
    clip_model = clip_model.to(device)
    ids = text_inputs["input_ids"].to(device)
    am = text_inputs.get("attention_mask", None)
    txt = d2r_clip.text_embeds(clip_model, ids, None if am is None else am.to(device))
    assert txt.shape[0] == len(captions), "text_inputs must hold one row per caption"

    with torch.no_grad():
        vision = _vision_for(clip_model, device.index, clip_batch_size)
        if stream_args is not None and len(mine) > 0:
            # render -> preprocess -> encode -> score per chunk: only [n_views, K] scores outlive a chunk (the reference renders
            # everything first, clip_scoring.py:120-147; at 800x800 its 70 000-pose shopping grid would be 134 GB of frames)
            valid_poses_ngp, render_poses_ngp = stream_args
            local = torch.empty((n_views, len(mine)), dtype=torch.float32, device=device)
            chunk = min(int(getattr(renderer, "max_candidates_per_launch", vision.max_batch)), vision.max_batch)
            for v, s, e, frames, rects, bg_u8 in renderer.iter_render(valid_poses_ngp[mine], render_poses_ngp, render_cam_pose_idx, depths_gt,
                                                                      task_model.movable_masks, save=save_renders and world == 1, chunk=chunk):
                emb = vision.encode_images(frames, rot90=True, bg_u8=bg_u8, rects=rects)
                local[v, s:e] = vision.score(emb, txt, n_goal=n_goal)
            # one score per pose = mean (or max) over its L per-view scores
            local = local[0] if n_views == 1 else (local.mean(0) if multi_view == "mean" else local.max(0).values)
        elif renders is not None and renders.shape[0] > 0:
            local = score_renders(renders, vision, txt, n_goal=n_goal)
        else:
            local = torch.zeros(0, dtype=torch.float32, device=device)
        logits = gather_scores(local, valid_idxs.shape[0], world, rank, mode=shard) if world > 1 else local
        logits = logits.to('cpu')

    pose_scores = torch.zeros(pose_batch.shape[0])
    pose_scores[valid_idxs.cpu()] = logits
    render_idxs = torch.zeros(pose_scores.shape[0], dtype=torch.long)
    render_idxs[valid_idxs.cpu()] = torch.arange(valid_idxs.shape[0])

    if smoothing:
        print('Applying spatial smoothing...')
        with torch.no_grad():
            pose_scores = spatially_smooth_heatmap(pose_scores, sample_res)
        print('Done smoothing.')

    best_pose_idx = torch.argmax(pose_scores).item()
    best_pose = valid_poses[render_idxs[best_pose_idx]]
    j = int(render_idxs[best_pose_idx])
    best_u8 = None
    if renders is not None:
        best_u8 = renders[j]
    elif stream_args is not None and rank == 0:
        # frames were streamed: render the winner again (a candidate's frame does not depend on its batch, bit for bit)
        best_u8 = renderer.render(stream_args[0][j:j + 1], stream_args[1][:1], render_cam_pose_idx[:1],
                                  None if depths_gt is None else depths_gt[:1], task_model.movable_masks, save=False, return_tensor=True)[0]
    task_model.free_visual_models()
    if best_u8 is not None:
        from PIL import Image
        best_render = np.rot90(best_u8.cpu().numpy(), k=1, axes=(0, 1))     # first view of the best pose
        best_render = Image.fromarray(np.ascontiguousarray(best_render))
        best_render.save(os.path.join(data_dir, 'best_render.png'))
        if show_best:
            best_render.show()
    return best_pose.view(4, 4), pose_batch, pose_scores


score_poses = optimise_pose_grid   # north-star alias

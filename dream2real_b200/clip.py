"""CLIP image side on libd2r_b200: preprocessing + vision tower + logits, fed with the weights of
a HuggingFace `CLIPModel` (what reference clip_scoring.py:150 loads).  The text tower runs once
per query through the caller's HF model in PyTorch.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # transformers/utils/constants.py
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

# (image_size, patch, hidden, heads, layers, mlp, proj) of the two configurations on the path
CLIP_CONFIGS = {
    "ViT-B/32": dict(image_size=224, patch_size=32, hidden=768, heads=12, layers=12, mlp=3072, proj=512,
                     text=dict(hidden=512, heads=8, layers=12, mlp=2048)),
    "ViT-L/14-336": dict(image_size=336, patch_size=14, hidden=1024, heads=16, layers=24, mlp=4096, proj=768,
                         text=dict(hidden=768, heads=12, layers=12, mlp=3072)),
}


def make_hf_clip(name: str = "ViT-B/32", seed: int = 1234, vocab_size: int = 49408):
    """Random-init HF CLIPModel of a named architecture (no network: weights are synthetic)."""
    import torch
    from transformers import CLIPConfig, CLIPModel, CLIPTextConfig, CLIPVisionConfig
    c = CLIP_CONFIGS[name]
    v = CLIPVisionConfig(hidden_size=c["hidden"], intermediate_size=c["mlp"], num_hidden_layers=c["layers"],
                         num_attention_heads=c["heads"], image_size=c["image_size"], patch_size=c["patch_size"],
                         projection_dim=c["proj"])
    t = CLIPTextConfig(hidden_size=c["text"]["hidden"], intermediate_size=c["text"]["mlp"], num_hidden_layers=c["text"]["layers"],
                       num_attention_heads=c["text"]["heads"], projection_dim=c["proj"], vocab_size=vocab_size)
    cfg = CLIPConfig(text_config=t.to_dict(), vision_config=v.to_dict(), projection_dim=c["proj"])
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = CLIPModel(cfg).eval()
    # default init leaves biases at 0 and LayerNorms at identity; perturb so every term is exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
            elif "layer_norm" in k or "layernorm" in k or "layrnorm" in k:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    return model


def vision_weight_list(model) -> List["np.ndarray"]:
    """Weights in the order include/d2r_b200.h documents for d2r_clip_load."""
    sd = {k: v.detach().float().cpu().contiguous().numpy() for k, v in model.state_dict().items()
          if k.startswith("vision_model.") or k.startswith("visual_projection.")}
    L = model.config.vision_config.num_hidden_layers
    out = [sd["vision_model.embeddings.patch_embedding.weight"], sd["vision_model.embeddings.class_embedding"],
           sd["vision_model.embeddings.position_embedding.weight"], sd["vision_model.pre_layrnorm.weight"],
           sd["vision_model.pre_layrnorm.bias"]]
    for l in range(L):
        p = f"vision_model.encoder.layers.{l}."
        for n in ("layer_norm1.weight", "layer_norm1.bias", "self_attn.q_proj.weight", "self_attn.q_proj.bias",
                  "self_attn.k_proj.weight", "self_attn.k_proj.bias", "self_attn.v_proj.weight", "self_attn.v_proj.bias",
                  "self_attn.out_proj.weight", "self_attn.out_proj.bias", "layer_norm2.weight", "layer_norm2.bias",
                  "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"):
            out.append(sd[p + n])
    out += [sd["vision_model.post_layernorm.weight"], sd["vision_model.post_layernorm.bias"], sd["visual_projection.weight"]]
    return [np.ascontiguousarray(a, dtype=np.float32) for a in out]


class ClipVision:
    """Vision tower + preprocessing of one HF CLIPModel on one GPU."""

    def __init__(self, hf_model, max_batch: int = 512, device: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ClipVision needs a CUDA device (B200); there is no CPU fallback")
        vc = hf_model.config.vision_config
        if vc.hidden_act != "quick_gelu":
            raise RuntimeError("only quick_gelu CLIP towers are on this path")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.image_size, self.patch = int(vc.image_size), int(vc.patch_size)
        self.hidden, self.proj = int(vc.hidden_size), int(hf_model.config.projection_dim)
        self.max_batch = int(max_batch)
        self.logit_scale_exp = float(hf_model.logit_scale.detach().exp().item())
        cfg = N.ClipCfg()
        cfg.image_size, cfg.patch_size, cfg.hidden = self.image_size, self.patch, self.hidden
        cfg.heads, cfg.layers, cfg.mlp = int(vc.num_attention_heads), int(vc.num_hidden_layers), int(vc.intermediate_size)
        cfg.proj, cfg.ln_eps, cfg.max_batch = self.proj, float(vc.layer_norm_eps), self.max_batch
        ws = vision_weight_list(hf_model)
        ptrs = (C.c_void_p * len(ws))(*[w.ctypes.data for w in ws])
        self._h = C.c_void_p()
        N.check(N.lib().d2r_clip_load(C.byref(cfg), ptrs, len(ws), self.device, C.byref(self._h)), "clip_load")
        self.n_patches = (self.image_size // self.patch) ** 2
        self.kp = (3 * self.patch * self.patch + 63) // 64 * 64
        self._patches = torch.empty((self.max_batch * self.n_patches, self.kp), dtype=torch.float16,
                                    device=torch.device("cuda", self.device))

    def preprocess(self, images_u8, rot90: bool = True, want_pixels: bool = False, bg_u8=None, rects=None):
        """uint8 CUDA [K,H,W,3] -> (patch-major fp16 [K*np, kp], optional float32 pixel_values [K,3,R,R]).
        bg_u8 [H,W,3] + rects int32 [K,4] (from Testbed.render_composite_batch): frames equal bg_u8 outside their
        rectangle, so only the affected part of every resize is recomputed (bit-identical result)."""
        import torch
        K, H, W, _ = images_u8.shape
        assert images_u8.is_cuda and images_u8.dtype == torch.uint8 and images_u8.is_contiguous() and K <= self.max_batch
        if bg_u8 is not None and rects is not None and not want_pixels:
            assert bg_u8.is_cuda and bg_u8.dtype == torch.uint8 and tuple(bg_u8.shape) == (H, W, 3) and bg_u8.is_contiguous()
            assert rects.is_cuda and rects.dtype == torch.int32 and tuple(rects.shape) == (K, 4) and rects.is_contiguous()
            with torch.cuda.device(images_u8.device):
                N.check(N.lib().d2r_clip_preprocess_delta(images_u8.data_ptr(), K, H, W, 1 if rot90 else 0, self.image_size, self.patch,
                                                          N.f4(OPENAI_CLIP_MEAN), N.f4(OPENAI_CLIP_STD), bg_u8.data_ptr(), rects.data_ptr(),
                                                          self._patches.data_ptr(), N.stream_ptr()), "clip_preprocess_delta")
            return self._patches[: K * self.n_patches], None
        pix = torch.empty((K, 3, self.image_size, self.image_size), dtype=torch.float32, device=images_u8.device) if want_pixels else None
        with torch.cuda.device(images_u8.device):
            N.check(N.lib().d2r_clip_preprocess(images_u8.data_ptr(), K, H, W, 1 if rot90 else 0, self.image_size, self.patch,
                                                N.f4(OPENAI_CLIP_MEAN), N.f4(OPENAI_CLIP_STD), self._patches.data_ptr(),
                                                pix.data_ptr() if want_pixels else None, N.stream_ptr()), "clip_preprocess")
        return self._patches[: K * self.n_patches], pix

    def encode_patches(self, patches, K: int):
        import torch
        out = torch.empty((K, self.proj), dtype=torch.float32, device=patches.device)
        with torch.cuda.device(patches.device):
            N.check(N.lib().d2r_clip_encode(self._h, patches.data_ptr(), K, out.data_ptr(), N.stream_ptr()), "clip_encode")
        return out

    def encode_images(self, images_u8, rot90: bool = True, bg_u8=None, rects=None):
        """L2-normalised image embeddings [K, proj] for uint8 CUDA images, batched by max_batch."""
        import torch
        outs = []
        for s in range(0, images_u8.shape[0], self.max_batch):
            chunk = images_u8[s:s + self.max_batch]
            patches, _ = self.preprocess(chunk, rot90, bg_u8=bg_u8, rects=None if rects is None else rects[s:s + self.max_batch])
            outs.append(self.encode_patches(patches, chunk.shape[0]))
        return torch.cat(outs, 0)

    def score(self, img_embeds, txt_embeds, n_goal: int = 1, want_logits: bool = False):
        """goal logit / mean(normalising logits) (clip_scoring.py:187-203); txt_embeds [C, proj] normalised."""
        import torch
        K, D = img_embeds.shape
        Cn = txt_embeds.shape[0]
        txt = txt_embeds.to(device=img_embeds.device, dtype=torch.float32).contiguous()
        scores = torch.empty(K, dtype=torch.float32, device=img_embeds.device)
        logits = torch.empty((K, Cn), dtype=torch.float32, device=img_embeds.device) if want_logits else None
        with torch.cuda.device(img_embeds.device):
            N.check(N.lib().d2r_score(img_embeds.data_ptr(), txt.data_ptr(), K, Cn, D, self.logit_scale_exp, n_goal,
                                      scores.data_ptr(), logits.data_ptr() if want_logits else None, N.stream_ptr()), "score")
        return (scores, logits) if want_logits else scores

    def __del__(self):
        try:
            if self._h:
                N.lib().d2r_clip_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def text_embeds(hf_model, input_ids, attention_mask=None):
    """Normalised text embeddings through the caller's HF text tower (once per query)."""
    import torch
    with torch.no_grad():
        out = hf_model.text_model(input_ids=input_ids, attention_mask=attention_mask)
        pooled = out.pooler_output if hasattr(out, "pooler_output") else out[1]
        e = hf_model.text_projection(pooled)
        return e / e.norm(p=2, dim=-1, keepdim=True)

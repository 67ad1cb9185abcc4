"""TEST INFRASTRUCTURE — CPU fp32 restatement of the AutoencoderKL *decoder* the reference calls for colour
guidance and the final decode (models/region_diffusion_sdxl.py:856, 938; models/region_diffusion.py:157, 232).

The module itself is third-party (`diffusers==0.18.2` `AutoencoderKL`, environment.yaml:15) and is NOT under
/root/reference, so this is a restatement of its published architecture (post_quant_conv -> conv_in -> mid block
[resnet, single-head attention, resnet] -> 4 up blocks of 3 resnets (+ nearest x2 upsample + conv) ->
GroupNorm/SiLU/conv_out) over a diffusers-format state dict. PARITY UNPINNED: no fixture of the real module
exists in this container (diffusers is not installed, there are no weights); what IS pinned is that the product
decoder engine (rtti_b200.vae_guidance) and this restatement agree, forward and d/dz, on seeded weights.
Plain functional PyTorch so autograd gives the reference's `loss.backward()` path (sdxl.py:865).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may import it.
"""
import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.13025   # SDXL; SD1.5 uses 0.18215


def param_shapes(cfg: VAEConfig):
    S = OrderedDict()
    boc = cfg.block_out_channels

    def conv(name, cin, cout, k):
        S[name + ".weight"] = (cout, cin, k, k); S[name + ".bias"] = (cout,)

    def lin(name, cin, cout):
        S[name + ".weight"] = (cout, cin); S[name + ".bias"] = (cout,)

    def norm(name, c):
        S[name + ".weight"] = (c,); S[name + ".bias"] = (c,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    conv("decoder.conv_in", cfg.latent_channels, boc[-1], 3)
    resnet("decoder.mid_block.resnets.0", boc[-1], boc[-1])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", boc[-1])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{n}", boc[-1], boc[-1])
    resnet("decoder.mid_block.resnets.1", boc[-1], boc[-1])
    prev = boc[-1]
    for i, c in enumerate(reversed(boc)):
        for l in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{l}", prev if l == 0 else c, c)
        if i != len(boc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
        prev = c
    norm("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", boc[0], cfg.out_channels, 3)
    return S


def make_state_dict(cfg: VAEConfig, seed: int):
    """Seeded synthetic weights (numpy PCG64), same recipe as unet_oracle.make_state_dict."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for k, shp in sorted(param_shapes(cfg).items()):
        if k.endswith(".weight") and len(shp) >= 2:
            w = rng.standard_normal(shp, dtype=np.float32) / math.sqrt(int(np.prod(shp[1:])))
        elif k.endswith(".weight"):
            w = 1.0 + 0.1 * rng.standard_normal(shp, dtype=np.float32)
        else:
            w = 0.05 * rng.standard_normal(shp, dtype=np.float32)
        sd[k] = torch.from_numpy(w)
    return sd


def _resnet(sd, name, x, groups):
    h = F.silu(F.group_norm(x, groups, sd[name + ".norm1.weight"], sd[name + ".norm1.bias"], 1e-6))
    h = F.conv2d(h, sd[name + ".conv1.weight"], sd[name + ".conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, groups, sd[name + ".norm2.weight"], sd[name + ".norm2.bias"], 1e-6))
    h = F.conv2d(h, sd[name + ".conv2.weight"], sd[name + ".conv2.bias"], padding=1)
    if (name + ".conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[name + ".conv_shortcut.weight"], sd[name + ".conv_shortcut.bias"])
    return x + h


def _mid_attention(sd, name, x, groups):
    """Single-head spatial self-attention with a residual connection (the VAE `Attention` block)."""
    B, C, H, W = x.shape
    h = F.group_norm(x, groups, sd[name + ".group_norm.weight"], sd[name + ".group_norm.bias"], 1e-6)
    h = h.view(B, C, H * W).transpose(1, 2)
    q = F.linear(h, sd[name + ".to_q.weight"], sd[name + ".to_q.bias"])
    k = F.linear(h, sd[name + ".to_k.weight"], sd[name + ".to_k.bias"])
    v = F.linear(h, sd[name + ".to_v.weight"], sd[name + ".to_v.bias"])
    p = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * (C ** -0.5), dim=-1)
    o = F.linear(torch.bmm(p, v), sd[name + ".to_out.0.weight"], sd[name + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def decode(sd, cfg: VAEConfig, z):
    """latents [B, 4, h, w] (already divided by scaling_factor) -> image [B, 3, 8h, 8w]; differentiable in z."""
    g = cfg.norm_num_groups
    boc = cfg.block_out_channels
    h = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _resnet(sd, "decoder.mid_block.resnets.0", h, g)
    h = _mid_attention(sd, "decoder.mid_block.attentions.0", h, g)
    h = _resnet(sd, "decoder.mid_block.resnets.1", h, g)
    for i in range(len(boc)):
        for l in range(cfg.layers_per_block + 1):
            h = _resnet(sd, f"decoder.up_blocks.{i}.resnets.{l}", h, g)
        if i != len(boc) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    h = F.silu(F.group_norm(h, g, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], 1e-6))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)

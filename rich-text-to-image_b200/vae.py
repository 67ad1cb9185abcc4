"""AutoencoderKL decoder (latents -> image) used for the final decode and for colour guidance.

Third-party on the reference side (diffusers 0.18.2 `AutoencoderKL`, called at
models/region_diffusion_sdxl.py:856,938 and models/region_diffusion.py:157,232); SURVEY §8(f).1 lists its
acceleration as the first "next" row.  It stays plain PyTorch here (cuDNN convolutions, fp32 with TF32
tensor cores, channels_last), with parameter names of the diffusers checkpoint so real weights load.
The weights are frozen: colour guidance needs d loss / d latents only, so autograd skips the weight
gradients the reference computes and never uses (models/region_diffusion_sdxl.py:865).
"""
import math
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215

    @staticmethod
    def sd15():
        return VAEConfig()

    @staticmethod
    def sdxl():
        return VAEConfig(scaling_factor=0.13025)


class _Resnet(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class _MidAttention(nn.Module):
    """Single-head spatial self-attention of the VAE mid block."""

    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o).transpose(1, 2).reshape(B, C, H, W)
        return x + o


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        if upsample:
            self.upsamplers = nn.ModuleList([nn.Module()])
            self.upsamplers[0].conv = nn.Conv2d(cout, cout, 3, padding=1)
        else:
            self.upsamplers = None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc = cfg.block_out_channels
        g = cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[-1], 3, padding=1)
        self.mid_block = nn.Module()
        self.mid_block.resnets = nn.ModuleList([_Resnet(boc[-1], boc[-1], g), _Resnet(boc[-1], boc[-1], g)])
        self.mid_block.attentions = nn.ModuleList([_MidAttention(boc[-1], g)])
        rev = list(reversed(boc))
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, c in enumerate(rev):
            self.up_blocks.append(_UpBlock(prev, c, cfg.layers_per_block + 1, g, upsample=i != len(rev) - 1))
            prev = c
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid_block.resnets[0](h)
        h = self.mid_block.attentions[0](h)
        h = self.mid_block.resnets[1](h)
        for b in self.up_blocks:
            h = b(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKLDecoder(nn.Module):
    """`decode(z).sample` interface of the reference's VAE; fp32 (the SDXL VAE overflows in fp16,
    models/region_diffusion_sdxl.py:916-917)."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.config = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = Decoder(cfg)

    def finalize(self, device="cuda"):
        self.to(device=device, dtype=torch.float32, memory_format=torch.channels_last)
        self.requires_grad_(False)
        return self.eval()

    def init_synthetic(self, seed=0):
        g = torch.Generator().manual_seed(seed)
        for name, p in self.named_parameters():
            if p.dim() >= 2:
                p.data.copy_(torch.randn(p.shape, generator=g) / math.sqrt(p[0].numel()))
            elif name.endswith("weight"):
                p.data.fill_(1.0)
            else:
                p.data.zero_()
        return self

    def decode_tensor(self, z):
        z = z.contiguous(memory_format=torch.channels_last)
        return self.decoder(self.post_quant_conv(z))

    def decode(self, z, return_dict=True):
        out = self.decode_tensor(z)

        class _O:
            sample = out
        return _O() if return_dict else (out,)

"""Seeded synthetic inputs shared by the golden-fixture generator (oracle/gen_golden.py keeps an identical
copy of these formulas) and the tests. Pure torch-CPU, deterministic for a given torch build."""
import torch


def synth_inputs(cross_attention_dim, pooled_dim, n_prompts, latent, seed):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(1, 4, latent, latent, generator=g)
    ctx = torch.randn(n_prompts + 1, 77, cross_attention_dim, generator=g)
    out = {"latents": lat, "ctx": ctx}
    if pooled_dim:
        out["text_embeds"] = torch.randn(n_prompts + 1, pooled_dim, generator=g)
        s = float(latent * 8)
        out["time_ids"] = torch.tensor([[s, s, 0.0, 0.0, s, s]])
    logits = torch.randn(n_prompts, 1, 8, 8, generator=g)
    up = torch.nn.functional.interpolate(logits, (latent, latent), mode="bicubic", align_corners=False)
    m = torch.softmax(up * 3.0, dim=0)
    out["masks"] = [m[i:i + 1].repeat(1, 4, 1, 1) for i in range(n_prompts)]
    return out


def font_sizes():
    return {"word_pos": torch.LongTensor([2, 5, 9]), "font_size": torch.FloatTensor([2.0, 0.5, -1.5])}


def color_dict(masks, latent, weight):
    up = torch.nn.functional.interpolate(masks[0], (latent * 8, latent * 8), mode="bicubic", antialias=True).clamp(0, 1)
    return {"target_RGB": [torch.tensor([0.99, 0.42, 0.62]).reshape(1, 3, 1, 1)], "guidance_start_step": 999,
            "color_guidance_weight": weight, "color_obj_atten": [up], "color_obj_atten_all": masks[0].clone()}


class TinyVAE:
    """The stand-in decoder of the fixtures: fixed 1x1 conv 4->3 then nearest x8 (differentiable)."""

    def __init__(self, device="cpu", dtype=torch.float32):
        g = torch.Generator().manual_seed(77)
        self.w = (torch.randn(3, 4, 1, 1, generator=g) * 0.5).to(device, dtype)

        class _C:
            scaling_factor = 0.13025
        self.config = _C()

    def decode_tensor(self, z):
        return torch.nn.functional.interpolate(torch.nn.functional.conv2d(z, self.w.to(z.dtype)), scale_factor=8.0, mode="nearest")

    def __call__(self, z):
        return self.decode_tensor(z)


def synth_maps(seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(32.0), torch.arange(32.0), indexing="ij")
    blob = ((yy >= 16).long() * 2 + (xx >= 16).long()).reshape(-1)
    selfm, crossm = {}, {}
    for li in range(3):
        same = (blob[:, None] == blob[None, :]).float()
        a = same * 1.0 + 0.05 * torch.rand(1024, 1024, generator=g)
        a = a / a.sum(-1, keepdim=True)
        selfm[f"l{li}.attn1"] = a[None]
    selfm["small.attn1"] = torch.rand(1, 256, 256, generator=g)
    for li, r in enumerate((32, 16)):
        yy2, xx2 = torch.meshgrid(torch.arange(float(r)), torch.arange(float(r)), indexing="ij")
        q = ((yy2 >= r // 2).long() * 2 + (xx2 >= r // 2).long()).reshape(-1)
        c = 0.01 * torch.rand(1, r * r, 77, generator=g)
        c[0, q == 0, 3] += 0.6
        c[0, q == 3, 7] += 0.5
        c[0, q == 3, 8] += 0.4
        crossm[f"c{li}.attn2"] = c
    return selfm, crossm

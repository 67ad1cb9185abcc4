"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's region-diffusion sampling loops.

Follows models/region_diffusion_sdxl.py:772-914 (SDXL rich / plain loops), models/region_diffusion.py:86-225
(SD1.5 rich / plain loops) and the four hook managers of each file, with the hooks expressed as
oracle.unet_oracle.AttnControl callbacks. Pinned against the unmodified reference loops driven
through oracle/ref_shim.py (tests/test_oracle_vs_reference.py) and by tests/golden/.
"""
import torch

from . import unet_oracle as uo

# utils/attention_utils.py:12-67 (the capture allow-lists; data, restated verbatim by necessity)
SelfAttentionLayers = [f"{b}.transformer_blocks.0.attn1" for b in (
    "down_blocks.0.attentions.0", "down_blocks.0.attentions.1", "down_blocks.1.attentions.0",
    "down_blocks.1.attentions.1", "down_blocks.2.attentions.0", "down_blocks.2.attentions.1",
    "mid_block.attentions.0", "up_blocks.1.attentions.0", "up_blocks.1.attentions.1", "up_blocks.1.attentions.2",
    "up_blocks.2.attentions.0", "up_blocks.2.attentions.1", "up_blocks.2.attentions.2", "up_blocks.3.attentions.0",
    "up_blocks.3.attentions.1", "up_blocks.3.attentions.2")]
CrossAttentionLayers = [f"{b}.transformer_blocks.0.attn2" for b in (
    "down_blocks.1.attentions.0", "down_blocks.2.attentions.0", "down_blocks.2.attentions.1",
    "mid_block.attentions.0", "up_blocks.1.attentions.0", "up_blocks.1.attentions.1", "up_blocks.1.attentions.2",
    "up_blocks.2.attentions.1")]
CrossAttentionLayers_XL = (
    [f"down_blocks.2.attentions.1.transformer_blocks.{i}.attn2" for i in (3, 4)]
    + [f"mid_block.attentions.0.transformer_blocks.{i}.attn2" for i in (0, 1, 2, 3)]
    + [f"up_blocks.0.attentions.0.transformer_blocks.{i}.attn2" for i in (1, 2, 3, 4, 5, 6, 7)]
    + ["up_blocks.1.attentions.0.transformer_blocks.0.attn2"])


class TokenMapCapture(uo.AttnControl):
    """register_tokenmap_hooks: region_diffusion_sdxl.py:959-1009 (xl=True) / region_diffusion.py:397-443.
    Captures the head-averaged map of batch row 1 from the 11th call of each module on. The SD1.5 variant
    reproduces the reference's `name in crossattn_maps` test in the self-attention branch (:423), which makes
    the self maps overwrite instead of accumulate."""

    def __init__(self, xl, cross_layers=None, self_layers=None, start_after=10):
        self.xl = xl
        self.cross_layers = cross_layers if cross_layers is not None else (CrossAttentionLayers_XL if xl else CrossAttentionLayers)
        self.self_layers = self_layers if self_layers is not None else SelfAttentionLayers
        self.start_after = start_after
        self.selfattn_maps, self.crossattn_maps, self.n_maps = {}, {}, {}

    def post_attn(self, name, probs_avg, probs):
        self.n_maps[name] = self.n_maps.get(name, 0) + 1
        if "attn2" in name:
            if name in self.cross_layers and self.n_maps[name] > self.start_after:
                if name in self.crossattn_maps:
                    self.crossattn_maps[name] = self.crossattn_maps[name] + probs_avg[1:2]
                else:
                    self.crossattn_maps[name] = probs_avg[1:2].clone()
        else:
            take = self.n_maps[name] > self.start_after and (self.xl or name in self.self_layers)
            if take:
                accumulate = (name in self.selfattn_maps) if self.xl else (name in self.crossattn_maps)
                if accumulate:
                    self.selfattn_maps[name] = self.selfattn_maps[name] + probs_avg[1:2]
                else:
                    self.selfattn_maps[name] = probs_avg[1:2].clone()


class FontSizeControl(uo.AttnControl):
    """register_fontsize_hooks: region_diffusion_sdxl.py:1112-1140 — attn_weights to every attn2."""

    def __init__(self, text_format_dict):
        wp, fs = text_format_dict.get("word_pos"), text_format_dict.get("font_size")
        self.attn_weights = {"word_pos": wp, "font_size": fs} if (wp is not None and fs is not None) else None

    def pre_attn(self, name):
        if self.attn_weights is not None and "attn2" in name:
            return None, self.attn_weights
        return None, None


class SelfAttnStore(uo.AttnControl):
    """register_selfattn_hooks: region_diffusion_sdxl.py:1064-1110 — keep P of every attn1 and the hidden
    feature of up_blocks.1.resnets.1 of the reference pass (only on feature-injection steps)."""

    def __init__(self, active):
        self.active = active
        self.store = {}

    def post_attn(self, name, probs_avg, probs):
        if self.active and "attn2" not in name:
            self.store[name] = probs

    def post_resnet(self, name, hidden):
        if self.active and name == "up_blocks.1.resnets.1":
            self.store[name] = hidden


class ReplaceControl(uo.AttnControl):
    """register_replacement_hooks: region_diffusion_sdxl.py:1018-1061."""

    def __init__(self, active, store):
        self.active, self.store = active, store

    def pre_attn(self, name):
        if self.active and "attn1" in name:
            return self.store[name], None
        return None, None

    def pre_resnet(self, name):
        if self.active and name == "up_blocks.1.resnets.1":
            return self.store[name]
        return None


def predict_x0(alphas_cumprod, x_t, eps_t, t):
    """region_diffusion_sdxl.py:955-957 / region_diffusion.py:176-178."""
    a = alphas_cumprod[int(t)]
    return (x_t - eps_t * torch.sqrt(1 - a)) / torch.sqrt(a)


def color_guidance(latents, noise_pred, t, alphas_cumprod, vae_decode, scaling_factor, text_format_dict, xl=True):
    """region_diffusion_sdxl.py:849-867 / region_diffusion.py:151-168. vae_decode: differentiable latents->image."""
    with torch.enable_grad():
        latents = latents.detach().requires_grad_(True)
        x0 = predict_x0(alphas_cumprod, latents, noise_pred, t)
        imgs = vae_decode(x0 / scaling_factor if xl else (1 / scaling_factor) * x0)
        imgs = (imgs / 2 + 0.5).clamp(0, 1)
        loss_total = 0.0
        for attn_map, rgb_val in zip(text_format_dict["color_obj_atten"], text_format_dict["target_RGB"]):
            avg_rgb = (imgs * attn_map[:, 0]).sum(2).sum(2) / attn_map[:, 0].sum()
            loss_total = loss_total + torch.nn.functional.mse_loss(avg_rgb, rgb_val[:, :, 0, 0]) * 100
        loss_total.backward()
    return (latents - latents.grad * text_format_dict["color_guidance_weight"] * text_format_dict["color_obj_atten_all"]).detach().clone()


def rich_text_loop(unet, scheduler, text_embeddings, masks, latents, num_inference_steps, guidance_scale, xl,
                   added_cond=None, use_guidance=False, text_format_dict=None, inject_selfattn=0.0,
                   inject_background=0.0, vae_decode=None, scaling_factor=0.18215, trace=None):
    """The region loop. `unet(sample, t, ctx, added, ctrl)` -> eps; text_embeddings = [uncond, region_1.., base].
    xl=True : models/region_diffusion_sdxl.py:772-878     xl=False: models/region_diffusion.py:86-174
    `added_cond` (xl): dict(text_embeds [N+1,P], time_ids [*,6]); per-pass rows as in :787-821."""
    tfd = text_format_dict or {}
    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    inject = inject_selfattn > 0 or inject_background > 0
    latents_reference = latents.clone() if inject else None
    n_t = len(timesteps)

    def added(rows):
        if added_cond is None:
            return None
        return {"text_embeds": added_cond["text_embeds"][rows], "time_ids": added_cond["time_ids"][:1]}

    last = text_embeddings.shape[0] - 1
    for i, t in enumerate(timesteps):
        feat_inject_step = bool(t > (1 - inject_selfattn) * 1000)
        if xl:
            background_inject_step = i < inject_background * n_t
        else:
            background_inject_step = (i == int(inject_background * n_t)) and inject_background > 0
        with torch.no_grad():
            x_in = scheduler.scale_model_input(latents, t) if xl else latents
            eps_u = unet(x_in, t, text_embeddings[:1], added(slice(0, 1)), None)
            eps_text_cur = unet(x_in, t, text_embeddings[-1:], added(slice(last, last + 1)), FontSizeControl(tfd))
            if inject:
                xr_in = scheduler.scale_model_input(latents_reference, t) if xl else latents_reference
                eps_u_ref = unet(xr_in, t, text_embeddings[:1], added(slice(0, 1)), None)
                store = SelfAttnStore(feat_inject_step)
                eps_t_ref = unet(xr_in, t, text_embeddings[-1:], added(slice(last, last + 1)), store)
            noise_pred_uncond = eps_u * masks[-1]
            noise_pred_text = eps_text_cur * masks[-1]
            for j, mask in enumerate(masks[:-1]):
                ctrl = ReplaceControl(feat_inject_step, store.store) if inject else None
                eps_j = unet(x_in, t, text_embeddings[j + 1:j + 2], added(slice(j + 1, j + 2)), ctrl)
                noise_pred_uncond = noise_pred_uncond + eps_u * mask
                noise_pred_text = noise_pred_text + eps_j * mask
            noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
            joint = (inject_selfattn > 0 or background_inject_step > 0) if xl else inject
            if joint:
                noise_pred_refer = eps_u_ref + guidance_scale * (eps_t_ref - eps_u_ref)
                both = scheduler.step(torch.cat([noise_pred, noise_pred_refer]), t,
                                      torch.cat([latents, latents_reference]))["prev_sample"]
                latents, latents_reference = torch.chunk(both, 2, dim=0)
            else:
                latents = scheduler.step(noise_pred, t, latents)["prev_sample"]
        if use_guidance and bool(t < tfd["guidance_start_step"]):
            latents = color_guidance(latents, noise_pred, t, scheduler.alphas_cumprod, vae_decode, scaling_factor, tfd, xl)
        if xl:
            do_bg = (i == int(inject_background * n_t)) and inject_background > 0
        else:
            do_bg = background_inject_step
        if do_bg:
            latents = latents_reference * masks[-1] + latents * (1 - masks[-1])
        if trace is not None:
            trace.append({"latents": latents.detach().clone(), "noise_pred": noise_pred.detach().clone()})
    return latents


def plain_loop(unet, scheduler, text_embeddings, latents, num_inference_steps, guidance_scale, xl, added_cond=None,
               ctrl=None):
    """CFG loop with batch 2 = [uncond, cond]: region_diffusion_sdxl.py:879-914 / region_diffusion.py:180-225.
    `ctrl` is typically a TokenMapCapture."""
    scheduler.set_timesteps(num_inference_steps)
    for t in scheduler.timesteps:
        x = torch.cat([latents] * 2)
        if xl:
            x = scheduler.scale_model_input(x, t)
        with torch.no_grad():
            eps = unet(x, t, text_embeddings, added_cond, ctrl)
        eu, et = eps.chunk(2)
        noise_pred = eu + guidance_scale * (et - eu)
        latents = scheduler.step(noise_pred, t, latents)["prev_sample"]
    return latents


def make_unet_fn(sd, cfg):
    def fn(sample, t, ctx, added, ctrl):
        return uo.unet_forward(sd, cfg, sample, t, ctx, added, ctrl)
    return fn

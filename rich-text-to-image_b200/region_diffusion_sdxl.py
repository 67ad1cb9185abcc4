"""RegionDiffusionXL — B200-native drop-in for models/region_diffusion_sdxl.py of the reference.

Same public surface (`sample(...)`, `register_tokenmap_hooks / remove_tokenmap_hooks`, `.masks`,
`.selfattn_maps / .crossattn_maps / .n_maps`, `.unet .vae .scheduler .tokenizer`) and the same per-step
semantics (models/region_diffusion_sdxl.py:772-914), re-organised for the hardware:

  * the 2 + 2*inject + (N-1) UNet passes of a step (uncond, base+font-size, reference uncond, reference
    base, N-1 regions; :787-821) run as ONE batched UNet call — they share the timestep and, up to the
    reference latent, the input; the hook choreography becomes a RegionControl;
  * region blend + CFG + Euler update is one kernel (rtti_region_blend_cfg); colour-guidance loss fwd/bwd,
    guidance update, x0 prediction and background injection are kernels too;
  * with torch.distributed initialised the passes are sharded over the ranks (region_parallel.py) and the
    per-pass noise predictions are all-gathered before the (replicated, deterministic) blend.
"""
import math
from typing import List, Optional

import numpy as np
import torch

from . import ops, region_parallel, vae_guidance
from .attention_utils import CrossAttentionLayers_XL
from .schedulers import EulerDiscreteScheduler
from .unet import CrossKVCache, RegionControl, TokenMapAccumulator, UNet2DConditionModel, UNetConfig
from .vae import AutoencoderKLDecoder, VAEConfig


class StableDiffusionXLPipelineOutput(dict):
    def __init__(self, images):
        super().__init__(images=images)
        self.images = images


class RegionDiffusionXL:
    def __init__(self, load_path: str = "stabilityai/stable-diffusion-xl-base-1.0", device: str = "cuda",
                 force_zeros_for_empty_prompt: bool = True, unet=None, vae=None, scheduler=None,
                 text_encoders=None):
        """Either pass pre-built components (tests / synthetic benchmarks) or a local diffusers-format
        directory as `load_path` (models/region_diffusion_sdxl.py:87-137 downloads from the hub; there is
        no network here, so only local paths are supported)."""
        self.device = torch.device(device)
        torch.backends.cudnn.benchmark = True   # static shapes: let cuDNN pick its fastest conv algorithm once
        self.device_type = device
        if unet is None:
            from .loading import load_sdxl_components
            unet, vae, scheduler, text_encoders = load_sdxl_components(load_path, self.device)
        self.unet = unet
        self.vae = vae
        self.scheduler = scheduler or EulerDiscreteScheduler()
        self.text_encoders = text_encoders
        self.tokenizer = getattr(text_encoders, "tokenizer", None)
        self.tokenizer_2 = getattr(text_encoders, "tokenizer_2", None)
        self.force_zeros_for_empty_prompt = force_zeros_for_empty_prompt
        self.vae_scale_factor = 8
        self.default_sample_size = unet.config.sample_size
        self.masks = []
        self.attention_maps = None
        self.selfattn_maps = None
        self.crossattn_maps = None
        self.n_maps = None
        self._capture = None
        self.capture_all_resolutions = False
        self._exchanges = {}
        self.use_cuda_graphs = True  # replay the batched UNet pass of a step as one CUDA graph (launch-bound otherwise)
        self.profile_events = None   # dict -> CUDA-event pairs per phase of a step (bench.py breakdown)
        self.fused_exchange = True   # multi-GPU: fused peer-memory gather+blend kernel instead of NCCL all-gather
        self.stripe_guidance = True  # multi-GPU: colour guidance (VAE fwd+bwd) split by image rows over the ranks
        self._stripe_engines = {}
        self.graph_guidance = True   # replay decode -> colour loss -> decoder backward as one CUDA graph (launch-bound on >1 GPU)
        self._guidance_graphs = {}
        self.region_group = None     # torch.distributed group the passes of one image are sharded over (None = all ranks)
        self.remote_qk = True        # multi-GPU: pass D on one rank, its Q|K / feature pushed to the region-pass ranks
        self._remote = {}            # (instead of replicating D on every rank that owns a region pass)
        self.last_step_stats = {}

    @classmethod
    def from_synthetic(cls, unet_cfg: Optional[UNetConfig] = None, vae_cfg: Optional[VAEConfig] = None, seed=0,
                       device="cuda", with_vae=True):
        """Random-weight model of the right architecture (benchmarks / tests; no checkpoints available)."""
        with torch.device(device):  # build directly on the GPU: CPU default-init of 2.6 B parameters is slow
            unet = UNet2DConditionModel(unet_cfg or UNetConfig.sdxl())
        unet.finalize(device).init_synthetic(seed)
        vae = None
        if with_vae:
            vae = AutoencoderKLDecoder(vae_cfg or VAEConfig.sdxl()).init_synthetic(seed + 1).finalize(device)
        return cls(device=device, unet=unet, vae=vae, scheduler=EulerDiscreteScheduler())

    # ------------------------------------------------------------------ token-map capture API
    def register_tokenmap_hooks(self):
        """models/region_diffusion_sdxl.py:959-1009 — here: arm the on-device accumulators."""
        res = None if self.capture_all_resolutions else (32,)
        self._capture = TokenMapAccumulator(CrossAttentionLayers_XL, self_layers=None, start_after=10,
                                            sd_overwrite_bug=False, self_resolutions=res)
        self.selfattn_maps = self._capture.selfattn_maps
        self.crossattn_maps = self._capture.crossattn_maps
        self.n_maps = self._capture.n_maps

    def remove_tokenmap_hooks(self):
        self._capture = None
        self.selfattn_maps = None
        self.crossattn_maps = None
        self.n_maps = None

    # ------------------------------------------------------------------ helpers
    def encode_prompt(self, prompt, negative_prompt):
        if self.text_encoders is None:
            raise RuntimeError("no text encoders loaded: pass prompt_embeds / pooled_prompt_embeds explicitly")
        return self.text_encoders.encode(prompt, negative_prompt, self.device, self.force_zeros_for_empty_prompt)

    def prepare_latents(self, height, width, generator=None, latents=None):
        shape = (1, self.unet.config.in_channels, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=self.device, dtype=torch.float16)
        else:
            latents = latents.to(self.device, torch.float16)
        return latents * self.scheduler.init_noise_sigma

    def predict_x0(self, x_t, eps_t, t):
        """models/region_diffusion_sdxl.py:955-957 (alphas_cumprod[int(t)] with the post-step latents)."""
        alpha = float(self.scheduler.alphas_cumprod[int(float(t))])
        return ops.predict_x0(x_t.contiguous(), eps_t.contiguous(), alpha), alpha

    def _color_guidance(self, latents, noise_pred, t, tfd):
        """models/region_diffusion_sdxl.py:849-867 with the clamp / masked mean / MSE forward+backward in
        rtti_color_loss_fwd_bwd and the VAE (third-party) in PyTorch autograd with frozen weights."""
        x0, alpha = self.predict_x0(latents, noise_pred, t)
        sf = self.vae.config.scaling_factor
        # the reference pairs maps and targets with zip() (sdxl.py:857 / region_diffusion.py:159): sample.py hands over
        # R colour maps + the background map but only R target colours, and the background map is dropped
        n_col = min(len(tfd["color_obj_atten"]), len(tfd["target_RGB"]))
        masks = torch.stack([m[0, 0].to(self.device, torch.float32) for m in tfd["color_obj_atten"][:n_col]]).contiguous()
        tgt = torch.stack([r.reshape(3).to(self.device, torch.float32) for r in tfd["target_RGB"][:n_col]]).contiguous()

        def grad_image(img):
            loss, g = ops.color_loss_fwd_bwd(img[0].contiguous(), masks, tgt)
            self.last_step_stats["color_loss"] = loss
            return g[None]

        z = x0.float() / sf
        engine = self._stripe_engine(x0)
        if self.graph_guidance and isinstance(self.vae, AutoencoderKLDecoder):
            eng = engine if engine is not None else vae_guidance.default_engine(self.vae)
            gkey = (id(eng), tuple(z.shape), tuple(masks.shape))
            gg = self._guidance_graphs.get(gkey)
            if gg is None:
                gg = self._guidance_graphs[gkey] = vae_guidance.GuidanceGraph(eng)
            self.last_step_stats["color_loss"], grad_lat = gg(z.contiguous(), masks, tgt)
            grad_lat = grad_lat / (sf * math.sqrt(alpha))
        else:
            grad_lat = vae_guidance.image_and_latent_grad(self.vae, z, grad_image, engine=engine) / (sf * math.sqrt(alpha))
        atten_all = tfd["color_obj_atten_all"].to(self.device, torch.float32).expand_as(grad_lat).contiguous()
        return ops.latent_guidance_update(latents.contiguous(), grad_lat.contiguous(), atten_all,
                                          float(tfd["color_guidance_weight"]))

    def _stripe_engine(self, x0):
        """Stripe-parallel VAE engine (stripe_parallel.py) when running on >1 GPU, else None (single-GPU engine)."""
        import torch.distributed as dist
        if not (self.stripe_guidance and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return None
        from .vae import AutoencoderKLDecoder
        h, w = int(x0.shape[2]), int(x0.shape[3])
        if not isinstance(self.vae, AutoencoderKLDecoder) or x0.shape[0] != 1 or h % dist.get_world_size() != 0:
            return None
        if (h, w) not in self._stripe_engines:   # symmetric arena allocated once per latent shape
            try:
                from .stripe_parallel import StripedDecoderFwdBwd
                self._stripe_engines[(h, w)] = StripedDecoderFwdBwd(self.vae, h, w, self.device)
            except Exception as e:               # no peer-mappable memory: replicated guidance
                import warnings
                warnings.warn(f"rtti_b200: stripe-parallel colour guidance unavailable ({e!r}); running it replicated")
                self.stripe_guidance = False
                return None
        return self._stripe_engines[(h, w)]

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    def sample(self, prompt=None, height: Optional[int] = None, width: Optional[int] = None,
               num_inference_steps: int = 50, guidance_scale: float = 5.0, negative_prompt=None,
               num_images_per_prompt: int = 1, eta: float = 0.0, generator=None, latents=None, prompt_embeds=None,
               negative_prompt_embeds=None, pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None,
               output_type: Optional[str] = "pil", return_dict: bool = True, callback=None, callback_steps: int = 1,
               cross_attention_kwargs=None, guidance_rescale: float = 0.0, original_size=None,
               crops_coords_top_left=(0, 0), target_size=None, use_guidance: bool = False,
               inject_selfattn: float = 0.0, inject_background: float = 0.0, text_format_dict: Optional[dict] = None,
               run_rich_text: bool = False):
        """Signature of models/region_diffusion_sdxl.py:556-587. `prompt` is the list of region prompts with the
        base prompt last (sample.py:107); embeddings may be passed instead of text."""
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        original_size = original_size or (height, width)
        target_size = target_size or (height, width)
        if guidance_rescale > 0.0 and run_rich_text:
            raise NotImplementedError  # as the reference, :826-829
        if prompt_embeds is None:
            prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds = \
                self.encode_prompt(prompt, negative_prompt)
        dev = self.device
        # [uncond, prompts...] like :760-763
        ctx = torch.cat([negative_prompt_embeds, prompt_embeds], 0).to(dev, torch.float16)
        pooled = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], 0).to(dev, torch.float16)
        time_ids = torch.tensor([list(original_size) + list(crops_coords_top_left) + list(target_size)],
                                dtype=torch.float32, device=dev)
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        timesteps = self.scheduler.timesteps
        latents = self.prepare_latents(height, width, generator, latents)

        if run_rich_text:
            latents = self._rich_text_loop(ctx, pooled, time_ids, latents, timesteps, guidance_scale, use_guidance,
                                           inject_selfattn, inject_background, text_format_dict or {}, callback,
                                           callback_steps)
        else:
            latents = self._plain_loop(ctx, pooled, time_ids, latents, timesteps, guidance_scale, callback, callback_steps)

        if output_type == "latent":
            return StableDiffusionXLPipelineOutput(images=latents)
        image = self.vae.decode_tensor(latents.float() / self.vae.config.scaling_factor)
        image = (image / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return StableDiffusionXLPipelineOutput(images=image)
        arr = (image.permute(0, 2, 3, 1).float().cpu().numpy() * 255).round().astype("uint8")
        if output_type == "np":
            return StableDiffusionXLPipelineOutput(images=arr)
        from PIL import Image
        return StableDiffusionXLPipelineOutput(images=[Image.fromarray(a) for a in arr])

    def _plain_loop(self, ctx, pooled, time_ids, latents, timesteps, guidance_scale, callback, callback_steps):
        """:879-914 — CFG batch [uncond, cond]; with capture armed the attention kernels accumulate the maps."""
        ctx2 = torch.cat([ctx[:1], ctx[-1:]])
        pooled2 = torch.cat([pooled[:1], pooled[-1:]])
        kv = CrossKVCache()
        ones = None
        for i, t in enumerate(timesteps):
            sigma = self.scheduler.sigma(t)
            x = (latents / math.sqrt(sigma * sigma + 1.0)).expand(2, -1, -1, -1)
            ctrl = RegionControl(capture=self._capture, capture_row=1, kv_cache=kv)
            eps = self.unet(x, t, ctx2, {"text_embeds": pooled2, "time_ids": time_ids}, ctrl)["sample"]
            n = eps[0].numel()
            if ones is None:
                ones = torch.ones(1, n, dtype=torch.float32, device=eps.device)
            _, latents = ops.region_blend_cfg(eps[0:1].contiguous(), [eps[1:2].contiguous()], ones, guidance_scale,
                                              latents=latents.contiguous(), dt_sigma=self.scheduler.dt(t))
            if callback is not None and i % callback_steps == 0:
                callback(i, t, latents)
        return latents

    def build_pass_batch(self, n_regions, inject):
        """Order of the batched passes of one step and the context row each one uses.
        rows index [uncond, region_1..region_{N-1}, base] (ctx); `x_ref` marks the reference-latent passes."""
        last = n_regions  # ctx row of the base prompt
        passes = [dict(kind="A", ctx=0, ref=False), dict(kind="B", ctx=last, ref=False)]
        if inject:
            passes += [dict(kind="C", ctx=0, ref=True), dict(kind="D", ctx=last, ref=True)]
        for j in range(n_regions - 1):
            passes.append(dict(kind="E", ctx=j + 1, ref=False, region=j))
        return passes

    def prepare_rich_text(self, ctx, pooled, time_ids, latents, timesteps, guidance_scale, use_guidance,
                          inject_selfattn, inject_background, tfd):
        """Everything of :772-778 that is constant over the steps, as a state object for rich_text_step()."""
        dev = self.device
        N = len(self.masks)
        assert ctx.shape[0] == N + 1, "prompts must be [region_1..region_{N-1}, base] matching self.masks"
        inject = inject_selfattn > 0 or inject_background > 0
        st = type("RichTextState", (), {})()
        st.ctx, st.pooled, st.time_ids = ctx, pooled, time_ids
        st.latents = latents
        st.latents_ref = latents.clone() if inject else None
        st.timesteps, st.n_t = timesteps, len(timesteps)
        st.guidance_scale, st.use_guidance = guidance_scale, use_guidance
        st.inject, st.inject_selfattn, st.inject_background = inject, inject_selfattn, inject_background
        st.tfd = tfd
        st.N = N
        st.masks = torch.stack([m.to(dev, torch.float32).reshape(-1) for m in self.masks]).contiguous()  # [N, n] (:776)
        st.ones = torch.ones(1, latents[0].numel(), dtype=torch.float32, device=dev)
        st.passes = self.build_pass_batch(N, inject)
        st.kind = {p["kind"] + str(p.get("region", "")): k for k, p in enumerate(st.passes)}
        st.plan = region_parallel.RegionParallelPlan(st.passes, inject, group=self.region_group,
                                                     remote_qk=self.remote_qk and self.fused_exchange)
        word_pos, font_size = tfd.get("word_pos"), tfd.get("font_size")
        if word_pos is not None and font_size is not None:
            if int(word_pos.max()) >= ctx.shape[1] or int(word_pos.min()) < 0:   # the reference's advanced indexing raises here
                raise IndexError(f"word_pos {word_pos.tolist()} outside the {ctx.shape[1]} text tokens")
            st.word_pos = word_pos.to(dev, torch.int32).contiguous()
            st.font_size = font_size.to(dev, torch.float32).contiguous()
        else:
            st.word_pos = st.font_size = None
        st.kv_caches = {}
        st.graphs = {}
        st.noise_pred = None
        return st

    def _unet_pass(self, st, x, t, local, feat_inject_step):
        """The batched UNet call of one step: eager, or (use_cuda_graphs) captured once per batch composition
        and replayed — the pass is ~1400 kernel launches whose CPU launch cost exceeds their GPU time."""
        passes, plan = st.passes, st.plan
        rows = [passes[p]["ctx"] for p in local]
        inj = bool(feat_inject_step and st.inject)
        key = (tuple(local), inj)
        kvc = st.kv_caches.setdefault(tuple(local), CrossKVCache())

        rq = self._remote_qk(st, x) if inj else None
        role = plan.remote_role(local) if rq is not None else None

        def make_ctrl(remote=True):
            ctrl = RegionControl(kv_cache=kvc)
            if inj:
                src = plan.injection_sources(local)                                     # :1018-1061
                if src is not None:         # None: this rank's region passes take pass D's tensors from another rank
                    ctrl.qk_src = src
                    ctrl.feature_src = src
                    ikey = ("idx",) + key
                    if ikey not in st.graphs:   # built once, outside any capture
                        st.graphs[ikey] = torch.as_tensor(src, device=self.device)
                    ctrl.feature_idx = st.graphs[ikey]
                if rq is not None and remote:
                    ctrl.remote = rq.begin_pass(role)
            if st.word_pos is not None:
                ctrl.word_pos, ctrl.font_size = st.word_pos, st.font_size               # :792-797
                ctrl.fs_batch_mask = sum(1 << k for k, p in enumerate(local) if passes[p]["kind"] == "B")
            return ctrl

        def unet(x_, t_, ctx_, added_, remote=True):
            """`remote=False`: the eager warm-up before a capture — pass D's tensors are neither pushed nor awaited
            (region passes use their own Q, K; the result is discarded). Producer and consumers then execute the
            hand-off exactly once per step — in the replayed graph — which is what keeps the single-buffered receive
            regions safe and every rank's sequence base in step."""
            out = self.unet(x_, t_, ctx_, added_, make_ctrl(remote))["sample"]
            if rq is not None and remote:
                rq.end_pass()
            return out

        if not self.use_cuda_graphs:
            return unet(x, t, st.ctx[rows], {"text_embeds": st.pooled[rows], "time_ids": st.time_ids})
        g = st.graphs.get(key)
        if g is None:
            g = {"x": torch.empty_like(x), "t": torch.zeros(1, dtype=torch.float32, device=self.device),
                 "ctx": st.ctx[rows].contiguous(), "added": {"text_embeds": st.pooled[rows].contiguous(), "time_ids": st.time_ids}}
            g["x"].copy_(x)
            g["t"].fill_(float(t))
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up outside capture: fills the prompt K/V cache, sets func attributes
                unet(g["x"], g["t"], g["ctx"], g["added"], remote=False)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(graph, pool=st.graphs.get("pool")):
                g["out"] = unet(g["x"], g["t"], g["ctx"], g["added"])
            g["launches"] = ops.LAUNCHES - n0   # rtti kernels inside the graph (for the launch accounting)
            if "pool" not in st.graphs:   # the graphs of one sampling call share a memory pool
                st.graphs["pool"] = graph.pool()
            g["graph"] = graph
            st.graphs[key] = g
        g["x"].copy_(x)
        g["t"].fill_(float(t))
        g["graph"].replay()
        ops._count(g["launches"])
        return g["out"]

    def _remote_qk(self, st, x):
        """RemoteQK buffers for this latent shape (created collectively by every rank of the region group on the first
        feature-injection step), or None: single GPU, remote_qk off, no peer-mappable memory, or the plan replicates D."""
        if not (st.plan.remote_qk and st.plan.world > 1):
            return None
        key = (int(x.shape[2]), int(x.shape[3]))
        if key not in self._remote:
            try:
                self._remote[key] = region_parallel.RemoteQK(self.unet.injection_layout(*key), self.device, group=self.region_group)
            except Exception as e:   # no peer-mappable memory (the same on every rank): replicate pass D instead
                import warnings
                warnings.warn(f"rtti_b200: RemoteQK unavailable ({e!r}); pass D is replicated on the region-pass ranks")
                self._remote[key] = None
        if self._remote[key] is None:
            st.plan.remote_qk = False
            st.plan._cache.clear()
        return self._remote[key]

    def rich_text_step(self, st, i):
        """One iteration of the region loop, models/region_diffusion_sdxl.py:779-878."""
        t = st.timesteps[i]
        passes, kind, plan, N = st.passes, st.kind, st.plan, st.N
        feat_inject_step = bool(float(t) > (1 - st.inject_selfattn) * 1000)            # :782
        background_inject_step = i < st.inject_background * st.n_t                      # :783
        sigma = self.scheduler.sigma(t)
        scale = 1.0 / math.sqrt(sigma * sigma + 1.0)                                    # :784
        if feat_inject_step and st.inject:
            self._remote_qk(st, st.latents)   # collective on first use: every rank of the group, also those without passes
        local = plan.local_passes(feat_inject_step)
        pe = self.profile_events
        if pe is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        if local:
            x = torch.cat([(st.latents_ref if passes[p]["ref"] else st.latents) for p in local]) * scale
            eps_local = self._unet_pass(st, x, t, local, feat_inject_step)
        else:   # more ranks than passes on this step: this rank only takes part in the exchange
            eps_local = st.latents.new_empty((0,) + tuple(st.latents.shape[1:]))
            rq = self._remote_qk(st, st.latents) if (feat_inject_step and st.inject) else None
            if rq is not None:   # keep this rank's sequence base in step with the ranks that ran a pass
                rq.begin_pass(None)
                rq.end_pass()
        if pe is not None:
            ev[1].record()
        dt = self.scheduler.dt(t)
        step_ref = st.inject and (st.inject_selfattn > 0 or background_inject_step)                       # :830-841
        ex = None
        if plan.world > 1 and self.fused_exchange:
            xkey = (tuple(p["kind"] for p in passes), st.latents[0].numel())
            if xkey not in self._exchanges:   # symmetric buffers are allocated once per problem shape
                try:
                    self._exchanges[xkey] = region_parallel.PeerExchange(passes, st.latents[0].numel(), self.device, group=self.region_group)
                except Exception as e:        # no peer-mappable memory (e.g. GPUs without P2P): NCCL all-gather path
                    import warnings
                    warnings.warn(f"rtti_b200: symmetric peer memory unavailable ({e!r}); using the NCCL all-gather exchange")
                    self.fused_exchange = False
            ex = self._exchanges.get(xkey)
        if ex is not None:
            # fused all-gather + blend + CFG + Euler over NVLink peer memory (csrc/gather_blend.cu)
            _, owner = plan._plan(feat_inject_step)
            sid = ex.publish(eps_local, local, owner)
            st.noise_pred, st.latents, ref_out = ops.gather_blend_step(
                ex.slot_ptrs, ex.flag_ptrs, ex.rank, ex.slot_owner(owner), N, st.masks, st.guidance_scale,
                st.latents.contiguous(), st.latents_ref.contiguous() if step_ref else None, dt, sid)
            if step_ref:
                st.latents_ref = ref_out
        else:
            eps = plan.gather(eps_local, local, feat_inject_step)   # NCCL all-gather; identity on one GPU
            one = lambda name: eps[kind[name]:kind[name] + 1].contiguous()
            regions = [one(f"E{j}") for j in range(N - 1)] + [one("B")]
            st.noise_pred, st.latents = ops.region_blend_cfg(one("A"), regions, st.masks, st.guidance_scale,
                                                              latents=st.latents.contiguous(), dt_sigma=dt)   # :810-825, :845
            if step_ref:
                _, st.latents_ref = ops.region_blend_cfg(one("C"), [one("D")], st.ones, st.guidance_scale,
                                                         latents=st.latents_ref.contiguous(), dt_sigma=dt)
        if pe is not None:
            ev[2].record()
        if st.use_guidance and float(t) < st.tfd["guidance_start_step"]:                                  # :849
            torch.cuda.nvtx.range_push("guidance")
            st.latents = self._color_guidance(st.latents, st.noise_pred, t, st.tfd)
            torch.cuda.nvtx.range_pop()
        if i == int(st.inject_background * st.n_t) and st.inject_background > 0:                          # :870-872
            st.latents = ops.bg_inject_blend(st.latents.contiguous(), st.latents_ref.contiguous(), st.masks[-1].contiguous())
        if pe is not None:
            ev[3].record()
            pe.update(unet=(ev[0], ev[1]), exchange_blend=(ev[1], ev[2]), color_guidance=(ev[2], ev[3]))
        return st.latents

    def _rich_text_loop(self, ctx, pooled, time_ids, latents, timesteps, guidance_scale, use_guidance,
                        inject_selfattn, inject_background, tfd, callback, callback_steps):
        """:772-878."""
        st = self.prepare_rich_text(ctx, pooled, time_ids, latents, timesteps, guidance_scale, use_guidance,
                                    inject_selfattn, inject_background, tfd)
        for i, t in enumerate(timesteps):
            self.rich_text_step(st, i)
            if callback is not None and i % callback_steps == 0:
                callback(i, t, st.latents)
        for ex in self._exchanges.values():
            ex.check()
        for eng in self._stripe_engines.values():
            eng.arena.check()
        return st.latents

"""CPU, build container only: the oracle against the UNMODIFIED reference imported through oracle/ref_shim.py
on inputs other than the committed fixtures. Skipped where /root/reference does not exist (the GPU box)."""
import pytest
import torch

from oracle import ref_shim, unet_oracle as uo

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    import torchvision  # noqa: F401  (before the stub modules are installed)
    return ref_shim.import_reference()


@pytest.mark.parametrize("cfg_fn,seed", [(uo.tiny_sd_config, 3), (uo.tiny_xl_config, 4)])
def test_unet_forward_bit_exact(ref, cfg_fn, seed):
    cfg = cfg_fn()
    sd = uo.make_state_dict(cfg, seed)
    model = ref.unet_2d_condition.UNet2DConditionModel(**cfg.ref_kwargs())
    model.load_state_dict(sd)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == {k: tuple(v) for k, v in uo.param_shapes(cfg).items()}
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    added = None
    if cfg.addition_embed_type:
        pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
        added = {"text_embeds": torch.randn(2, pooled, generator=g), "time_ids": torch.tensor([[128.0, 128, 0, 0, 128, 128]] * 2)}
    with torch.no_grad():
        for t in (torch.tensor(999), torch.tensor(37.0, dtype=torch.float64)):
            yr = model(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added)["sample"]
            yo = uo.unet_forward(sd, cfg, x, t, ctx, added)
            assert torch.equal(yr, yo)


def test_full_size_parameter_inventories(ref):
    for cfg, nparams in ((uo.sd15_config(), 859.5e6), (uo.sdxl_config(), 2567.5e6)):
        with torch.device("meta"):
            model = ref.unet_2d_condition.UNet2DConditionModel(**cfg.ref_kwargs())
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        mine = {k: tuple(v) for k, v in uo.param_shapes(cfg).items()}
        assert shapes == mine
        assert abs(sum(torch.Size(s).numel() for s in mine.values()) - nparams) < 0.1e6


def test_attention_fontsize_and_injection(ref):
    torch.manual_seed(0)
    attn = ref.attention_processor.Attention(query_dim=64, cross_attention_dim=48, heads=2, dim_head=32)
    sd = {"a." + k: v for k, v in attn.state_dict().items()}
    hs, ctx = torch.randn(1, 32, 64), torch.randn(1, 77, 48)
    aw = {"word_pos": torch.LongTensor([1, 4, 4]), "font_size": torch.FloatTensor([3.0, -2.0, 0.25])}

    class C(uo.AttnControl):
        def pre_attn(self, name):
            return None, aw

        def post_attn(self, name, pavg, p):
            self.out = (pavg, p)

    c = C()
    with torch.no_grad():
        o_ref, (pavg_ref, p_ref) = attn(hs, None, aw, encoder_hidden_states=ctx)
        o = uo.attention(sd, "a", 2, hs, ctx, c)
    assert torch.allclose(o, o_ref, atol=1e-6) and torch.allclose(c.out[0], pavg_ref, atol=1e-7)
    assert torch.allclose(c.out[1], p_ref, atol=1e-7)


# --------------------------------------------------------------------------- host-side text preparation
class _Tok:
    def _tokenize(self, text):
        return text.lower().replace(",", " ,").split()


class _Model:
    tokenizer = _Tok()


_DELTAS = [
    {"ops": [{"insert": "a church "}, {"attributes": {"color": "#fd6c9e"}, "insert": "garden"},
             {"insert": " with "}, {"attributes": {"font": "slabo"}, "insert": "mountains"},
             {"attributes": {"size": "60px"}, "insert": " snowy"}, {"attributes": {"link": "a red sun"}, "insert": " sky"},
             {"insert": "\n"}]},
    {"ops": [{"attributes": {"font": "mirza"}, "insert": "a lake"}, {"attributes": {"font": "mirza"}, "insert": " at dawn"},
             {"insert": ", "}, {"attributes": {"color": "#00ff00", "size": "18px", "strike": True}, "insert": "reeds"},
             {"insert": " and a "}, {"attributes": {"color": "#a52a2a"}, "insert": "boat"}, {"insert": "\n"}]},
    {"ops": [{"insert": "a plain prompt without attributes\n"}]},
]


@pytest.mark.parametrize("delta", _DELTAS)
def test_richtext_utils_match_the_reference_functions(ref, delta):
    """rtti_b200.richtext_utils against utils/richtext_utils.py of the unmodified reference on the same Quill deltas:
    parse_json :74-136, get_region_diffusion_input :139-185, get_attention_control_input :188-209,
    get_gradient_guidance_input :212-234 — identical prompts, token ids, font sizes and target colours."""
    from rtti_b200 import richtext_utils as ru
    rr = ref.richtext_utils

    def same(a, b):
        if torch.is_tensor(a) or torch.is_tensor(b):
            return torch.is_tensor(a) and torch.is_tensor(b) and a.shape == b.shape and torch.allclose(a.float().cpu(), b.float().cpu())
        if isinstance(a, (list, tuple)):
            return isinstance(b, (list, tuple)) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return a == b

    out_r = rr.parse_json(delta)
    out_p = ru.parse_json(delta, device="cpu")
    assert len(out_r) == len(out_p) == 9
    for i, (a, b) in enumerate(zip(out_r, out_p)):
        assert same(a, b), f"parse_json output {i}: {a!r} vs {b!r}"
    base, styles, notes, note_t, cspans, cnames, crgbs, sizes, use_grad = out_r
    pr, idr, btr = rr.get_region_diffusion_input(_Model(), base, styles, notes, note_t, cspans, cnames)
    pp, idp, btp = ru.get_region_diffusion_input(_Model(), base, styles, notes, note_t, cspans, cnames)
    assert pr == pp and btr == btp and same(idr, idp)
    tr = rr.get_attention_control_input(_Model(), btr, sizes)
    tp = ru.get_attention_control_input(_Model(), btp, sizes, device="cpu")
    assert set(tr) == set(tp) and all(same(tr[k], tp[k]) for k in tr)
    tr2, cr = rr.get_gradient_guidance_input(_Model(), btr, cspans, crgbs, dict(tr), color_guidance_weight=0.5)
    tp2, cp = ru.get_gradient_guidance_input(_Model(), btp, cspans, out_p[6], dict(tp), color_guidance_weight=0.5)
    assert same(cr, cp) and set(tr2) == set(tp2)
    for k in tr2:
        assert same(tr2[k], tp2[k]), k


@pytest.mark.parametrize("n_prompts,steps,inject_selfattn,inject_background,use_guidance,with_fs,seed", [
    (4, 3, 0.4, 0.0, False, True, 71),      # more regions, self-attention / feature injection on the first step only
    (2, 3, 0.0, 0.4, True, False, 72),      # configs[3]-like: background injection only (the joint-stepping quirk), colour guidance
])
def test_xl_rich_loop_live_reference_other_settings(ref, n_prompts, steps, inject_selfattn, inject_background, use_guidance,
                                                    with_fs, seed):
    """RegionDiffusionXL.sample(run_rich_text=True) of the UNMODIFIED reference (models/region_diffusion_sdxl.py:772-878)
    against the oracle's rich_text_loop at settings the committed fixtures do not cover: other region counts, step
    counts, injection windows, with / without font sizes and colour guidance."""
    from oracle import gen_golden as gg, sampler_oracle as sam, schedulers_oracle as so
    from tests import synth
    if ref.region_diffusion_sdxl is None:
        pytest.skip(ref.region_diffusion_sdxl_error)
    cfg = uo.tiny_xl_config()
    S = 128   # the reference asserts a 64-wide injected feature map (sdxl.py:1090): 1024^2 images only
    inp = gg.synth_inputs(cfg, n_prompts, S, seed)
    ctx, te = inp["ctx"], inp["text_embeds"]
    m = gg.make_xl_sampler(ref, cfg, 5, (ctx[1:], ctx[:1], te[1:], te[:1]))
    m.masks = inp["masks"]
    tfd = gg.text_format(1, S, seed, with_fs=with_fs)
    if use_guidance:
        tfd.update(gg.color_dict(inp["masks"], S, weight=0.7))
    out = m.sample(["p"] * n_prompts, height=S * 8, width=S * 8, num_inference_steps=steps, guidance_scale=6.0,
                   negative_prompt=[""], latents=inp["latents"].clone(), output_type="latent", use_guidance=use_guidance,
                   inject_selfattn=inject_selfattn, inject_background=inject_background, text_format_dict=dict(tfd),
                   run_rich_text=True).images.detach()
    sd = uo.make_state_dict(cfg, 5)
    sch = so.EulerDiscreteSchedulerOracle()
    sch.set_timesteps(steps)
    added = {"text_embeds": te, "time_ids": inp["time_ids"]}
    lat = sam.rich_text_loop(sam.make_unet_fn(sd, cfg), sch, ctx, inp["masks"], inp["latents"].clone() * sch.init_noise_sigma,
                             steps, 6.0, xl=True, added_cond=added, use_guidance=use_guidance, text_format_dict=dict(tfd),
                             inject_selfattn=inject_selfattn, inject_background=inject_background,
                             vae_decode=synth.TinyVAE() if use_guidance else None, scaling_factor=0.13025)
    assert torch.isfinite(out).all()
    torch.testing.assert_close(lat, out, atol=5e-4, rtol=1e-4)

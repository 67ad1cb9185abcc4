"""Region-parallel parity check, run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py

Every rank runs the tiny SDXL-shaped rich-text loop (3 regions, injection, font sizes, colour guidance) with
(a) the fused peer-memory gather+blend kernel and (b) the NCCL all-gather baseline, and compares both with the
golden latents produced by the unmodified single-process reference (tests/golden/xl_loops.npz); then the
stripe-parallel colour-guidance engine is compared with the single-GPU one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import synth  # noqa: E402


def check_stripe_guidance(rank, world):
    """Stripe-parallel VAE forward + backward (stripe_parallel.py) against the single-GPU engine on the same
    inputs: a small decoder (2 calls: buffer re-use across calls) and the SDXL-size one (timed)."""
    from rtti_b200.stripe_parallel import StripedDecoderFwdBwd
    from rtti_b200.vae import AutoencoderKLDecoder, VAEConfig
    from rtti_b200.vae_guidance import DecoderFwdBwd
    ok = True
    for name, cfg, hw, reps in (("small", VAEConfig(block_out_channels=(32, 64, 128, 128)), 32, 2),
                                ("sdxl", VAEConfig.sdxl(), 128, 3)):
        vae = AutoencoderKLDecoder(cfg).init_synthetic(seed=5).finalize("cuda")
        ref_eng = DecoderFwdBwd(vae)
        eng = StripedDecoderFwdBwd(vae, hw, hw, torch.device("cuda"))
        gen = torch.Generator().manual_seed(11)
        for rep in range(reps):
            z = torch.randn(1, 4, hw, hw, generator=gen).cuda()
            wgt = torch.randn(1, 3, hw * 8, hw * 8, generator=gen).cuda()
            grad_fn = lambda img: torch.tanh(img) * wgt
            img_r = ref_eng.forward(z)
            g_r = ref_eng.backward(grad_fn(img_r))
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            img_s = eng.forward(z)
            g_s = eng.backward(grad_fn(img_s))
            e1.record()
            img_r = ref_eng.forward(z)
            g_r = ref_eng.backward(grad_fn(img_r))
            e2.record()
            torch.cuda.synchronize()
            eng.arena.check()
            e_img = float((img_s - img_r).abs().max() / img_r.abs().max())
            e_g = float((g_s - g_r).abs().max() / g_r.abs().max())
            gathered = [torch.empty_like(g_s) for _ in range(world)]
            dist.all_gather(gathered, g_s.contiguous())
            same = all(torch.equal(gathered[0], x) for x in gathered)
            good = e_img < 2e-3 and e_g < 5e-3 and same and bool(torch.isfinite(g_s).all())
            ok &= good
            if rank == 0:
                print(f"stripe guidance [{name}] world={world} rep={rep}: image rel err {e_img:.2e}, latent-grad rel err {e_g:.2e}, "
                      f"ranks bit-identical: {same}, striped {e0.elapsed_time(e1):.2f} ms vs replicated {e1.elapsed_time(e2):.2f} ms "
                      f"({'OK' if good else 'FAIL'})", flush=True)
        # the same evaluation as ONE replayed CUDA graph (vae_guidance.GuidanceGraph: call 1 eager, call 2 captures, then
        # replays; the exchange sequence numbers come from the device-side base words)
        from rtti_b200 import ops
        from rtti_b200.vae_guidance import GuidanceGraph
        gg = GuidanceGraph(eng)
        masks = torch.rand(1, hw * 8, hw * 8, generator=gen).cuda()
        tgt = torch.tensor([[0.9, 0.4, 0.6]], device="cuda")
        for rep in range(4):
            z = torch.randn(1, 4, hw, hw, generator=gen).cuda()
            img_r = ref_eng.forward(z)
            loss_r, gi = ops.color_loss_fwd_bwd(img_r[0].contiguous(), masks, tgt)
            g_r = ref_eng.backward(gi[None])
            loss_e, g_e = gg._run(z, masks, tgt)             # the same striped evaluation issued eagerly (every rank does)
            g_e, loss_e = g_e.clone(), float(loss_e)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss_s, g_s = gg(z, masks, tgt)
            e1.record()
            torch.cuda.synchronize()
            eng.arena.check()
            # replayed graph vs eager launches of the same striped engine: the same kernels on the same data -> equal up to
            # nothing (asserted <= 1e-6); striped vs single-GPU engine: TF32 convolutions on different tile shapes, and the
            # colour-loss gradient (a near-uniform image gradient) cancels heavily in the backward pass -> 3e-2 of the range
            e_ge = float((g_s - g_e).abs().max() / g_e.abs().max())
            e_g = float((g_s - g_r).abs().max() / g_r.abs().max())
            e_l = abs(float(loss_s) - float(loss_r)) / abs(float(loss_r))
            gathered = [torch.empty_like(g_s) for _ in range(world)]
            dist.all_gather(gathered, g_s.contiguous())
            same = all(torch.equal(gathered[0], x) for x in gathered)
            good = e_ge <= 1e-6 and float(loss_s) == loss_e and e_g < 3e-2 and e_l < 2e-3 and same and bool(torch.isfinite(g_s).all())
            ok &= good
            if rank == 0:
                mode = "eager" if rep == 0 else ("capture+replay" if rep == 1 else "replay")
                print(f"graphed guidance [{name}] world={world} call {rep} ({mode}): vs eager striped {e_ge:.1e}, vs single-GPU engine "
                      f"latent-grad rel err {e_g:.2e}, loss rel err {e_l:.2e}, ranks bit-identical: {same}, {e0.elapsed_time(e1):.2f} ms "
                      f"({'OK' if good else 'FAIL'})", flush=True)
        del gg, eng, ref_eng, vae
        torch.cuda.empty_cache()
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import unet_oracle as uo
    from rtti_b200.region_diffusion_sdxl import RegionDiffusionXL
    from rtti_b200.unet import UNet2DConditionModel, UNetConfig
    cfg = uo.tiny_xl_config()
    unet = UNet2DConditionModel(UNetConfig.from_dict(cfg.__dict__))
    unet.load_state_dict(uo.make_state_dict(cfg, 2))
    unet.finalize("cuda")
    g = np.load(os.path.join(ROOT, "tests", "golden", "xl_loops.npz"))
    S = 128
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    inp = synth.synth_inputs(cfg.cross_attention_dim, pooled, 3, S, 31)
    ctx, te = inp["ctx"].cuda(), inp["text_embeds"].cuda()
    tfd = synth.font_sizes()
    tfd.update(synth.color_dict(inp["masks"], S, 1.0))
    ok = True
    results = {}
    for fused, remote in ((True, True), (True, False), (False, False)):
        model = RegionDiffusionXL(device="cuda", unet=unet, vae=synth.TinyVAE("cuda"))
        model.fused_exchange, model.remote_qk = fused, remote
        model.masks = [m.cuda() for m in inp["masks"]]
        out = model.sample(height=S * 8, width=S * 8, num_inference_steps=4, guidance_scale=8.5, latents=inp["latents"].clone(),
                           prompt_embeds=ctx[1:], negative_prompt_embeds=ctx[:1], pooled_prompt_embeds=te[1:],
                           negative_pooled_prompt_embeds=te[:1], output_type="latent", run_rich_text=True, use_guidance=True,
                           inject_selfattn=0.5, inject_background=0.5, text_format_dict=tfd).images.float()
        ref = torch.from_numpy(g["rich_latents"]).cuda()
        err = (out - ref).abs()
        tol = 5e-3 * ref.abs().max() + 3e-2 * ref.abs()
        good = bool((err <= tol).all())
        # replicated blend must keep the ranks bit-identical
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out.contiguous())
        same = all(torch.equal(gathered[0], x) for x in gathered)
        results[fused] = out
        for rq in model._remote.values():
            if rq is not None and rq.error():
                good = False
                print(f"rank {rank}: RemoteQK wait timed out", flush=True)
        ok &= good and same
        if rank == 0:
            print(f"world={world} fused_exchange={fused} remote_qk={remote} (pass D {'on one rank, Q|K pushed' if remote else 'replicated'}): "
                  f"max err vs reference golden {err.max().item():.4f} "
                  f"({'OK' if good else 'FAIL'}), ranks bit-identical: {same}", flush=True)
    # fewer passes than ranks on the non-injection steps (world >= 4): some ranks only take part in the exchange
    inp2 = synth.synth_inputs(cfg.cross_attention_dim, pooled, 2, S, 32)
    ctx2, te2 = inp2["ctx"].cuda(), inp2["text_embeds"].cuda()
    tfd2 = synth.font_sizes()
    tfd2.update(synth.color_dict(inp2["masks"], S, 1.0))
    res2 = {}
    for fused in (True, False):
        model = RegionDiffusionXL(device="cuda", unet=unet, vae=synth.TinyVAE("cuda"))
        model.fused_exchange = fused
        model.masks = [m.cuda() for m in inp2["masks"]]
        res2[fused] = model.sample(height=S * 8, width=S * 8, num_inference_steps=4, guidance_scale=8.5,
                                   latents=inp2["latents"].clone(), prompt_embeds=ctx2[1:], negative_prompt_embeds=ctx2[:1],
                                   pooled_prompt_embeds=te2[1:], negative_pooled_prompt_embeds=te2[:1], output_type="latent",
                                   run_rich_text=True, use_guidance=True, inject_selfattn=0.5, inject_background=0.5,
                                   text_format_dict=tfd2).images.float()
        gathered = [torch.empty_like(res2[fused]) for _ in range(world)]
        dist.all_gather(gathered, res2[fused].contiguous())
        same = all(torch.equal(gathered[0], x) for x in gathered)
        fin = bool(torch.isfinite(res2[fused]).all())
        ok &= same and fin
        if rank == 0:
            print(f"world={world} 2-prompt run (3 passes on non-injection steps) fused_exchange={fused}: finite {fin}, "
                  f"ranks bit-identical: {same}", flush=True)
    d2 = (res2[True] - res2[False]).abs().max().item()
    ok &= d2 <= 2e-2 * float(res2[False].abs().max())
    if rank == 0:
        print(f"2-prompt run: fused vs NCCL path max |diff| = {d2:.3e}", flush=True)
    ok &= check_stripe_guidance(rank, world)
    if rank == 0:
        d = (results[True] - results[False]).abs().max().item()
        print(f"fused vs NCCL path max |diff| = {d:.3e}", flush=True)
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

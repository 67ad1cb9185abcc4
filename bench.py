"""bench.py — denoising steps/sec of the region-diffusion hot path on SDXL 1024^2 with 5 regions.

    python bench.py --gpus N --steps K --warmup W            # product arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # reference arm: CPU oracle port on the host cores

Workload (BASELINE.json configs[2]): SDXL UNet (2.57 B params, random weights — no
checkpoints in this environment), latents [1,4,128,128], N=5 region prompts (4 regions + base),
inject_selfattn=0.5, inject_background=0.5 -> 8 UNet passes per step (uncond, base+font sizes, reference
uncond/base, 4 regions) run as one batched call; region blend + CFG + Euler; colour guidance through the
fp32 SDXL VAE decoder (1 colour region, weight 1). One "step" = one iteration of
models/region_diffusion_sdxl.py:779-878. Synthetic data, seeded.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_REGIONS = 5
NUM_INFERENCE_STEPS = 41
GUIDANCE = 8.5
PASSES_PER_STEP = 2 + 2 + (N_REGIONS - 1)
UNET_PASS_GFLOP = 6761.2  # SURVEY §8d [probe], batch-1 SDXL UNet forward


WORKLOAD = ("SDXL 1024x1024 font-color example shape: 5 region prompts, color_guidance_weight=1, inject_selfattn=0.5, "
            "inject_background=0.5, 8 UNet passes/step, 41-step Euler schedule")


def bench_config():
    """`config` of the JSON line: identical for the product arm and the --impl reference arm."""
    return {"workload": WORKLOAD, "passes_per_step": PASSES_PER_STEP,
            "unet_tflop_per_step": PASSES_PER_STEP * UNET_PASS_GFLOP / 1e3,
            "l2": "inputs larger than L2: 5.1 GB of fp16 UNet weights stream every step",
            "vae": "SDXL AutoencoderKL decoder, random weights, fp32/TF32, fwd+bwd inside the step"}


def synth_workload(device):
    import torch
    g = torch.Generator().manual_seed(7)
    N = N_REGIONS
    h = w = 128
    ctx = torch.randn(N + 1, 77, 2048, generator=g)
    pooled = torch.randn(N + 1, 1280, generator=g)
    latents = torch.randn(1, 4, h, w, generator=g)
    logits = torch.randn(N, 1, 8, 8, generator=g)
    up = torch.nn.functional.interpolate(logits, (h, w), mode="bicubic", align_corners=False)
    m = torch.softmax(up * 3.0, dim=0)
    masks = [m[i:i + 1].repeat(1, 4, 1, 1) for i in range(N)]
    color_mask = torch.nn.functional.interpolate(masks[0], (h * 8, w * 8), mode="bicubic", antialias=True).clamp(0, 1)
    tfd = {"word_pos": torch.LongTensor([2, 5, 9]), "font_size": torch.FloatTensor([2.0, 0.5, -1.5]),
           "target_RGB": [torch.tensor([253, 108, 158.0]).reshape(1, 3, 1, 1) / 255.0], "guidance_start_step": 999,
           "color_guidance_weight": 1.0, "color_obj_atten": [color_mask], "color_obj_atten_all": masks[0].clone()}
    return dict(ctx=ctx, pooled=pooled, latents=latents, masks=masks, tfd=tfd)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def host_threads():
    """CPU threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(math.ceil(int(quota) / int(period)))))
    except Exception:
        pass
    return n


def cpu_oracle_pass_time(n_samples, warm):
    """Seconds per batch-1 SDXL UNet pass of the CPU oracle (fp32, all usable host threads), from a BOUNDED sample:
    conv_in + all three down blocks (128^2 resnets, four 64^2 and twenty 32^2 transformer layers: every level of the UNet,
    ~40 % of the pass) are executed and timed; the rest of the pass is extrapolated by FLOPs (counted on the sample with FlopCounterMode, whole pass =
    6761.2 GFLOP). Returns (seconds per pass, threads, seconds per sample, sample FLOP fraction)."""
    import torch
    from torch.utils.flop_counter import FlopCounterMode
    from oracle import unet_oracle as uo
    torch.set_num_threads(host_threads())
    cfg = uo.sdxl_config()
    g = torch.Generator().manual_seed(0)
    sd = {}
    for k, shp in uo.param_shapes(cfg).items():
        if not (k.startswith("conv_in") or k.startswith("time_embedding") or k.startswith("add_embedding")
                or k.startswith("down_blocks.")):
            continue
        if len(shp) >= 2:
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(float(torch.Size(shp[1:]).numel()))
        else:
            sd[k] = torch.ones(shp) if k.endswith("weight") else torch.zeros(shp)
    x = torch.randn(1, 4, 128, 128, generator=g)
    ctx = torch.randn(1, 77, 2048, generator=g)
    added = {"text_embeds": torch.randn(1, 1280, generator=g), "time_ids": torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]])}
    run = lambda: uo.unet_forward(sd, cfg, x, torch.tensor(981.0), ctx, added, stop_after_down_block=2)
    with torch.no_grad():
        with FlopCounterMode(display=False) as fc:
            run()
        frac = fc.get_total_flops() / (UNET_PASS_GFLOP * 1e9)
        times = []
        for i in range(warm + n_samples):
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    t_sample = sum(times) / len(times)
    return t_sample / frac, torch.get_num_threads(), t_sample, frac


def run_reference(args, rank):
    """Reference arm: the reference's own algorithm on the host cores. The reference is pure Python on top
    of `diffusers`, which is not installed and cannot be on the GPU box, so this is the CPU oracle port
    (oracle/unet_oracle.py, pinned against the unmodified reference by tests/golden). Each "step" times a
    bounded sample: ONE of the 8 batch-1 UNet passes of a step; steps/s = 1 / (8 * seconds per pass)."""
    if rank != 0:
        return
    per_pass, threads, t_sample, frac = cpu_oracle_pass_time(args.steps, min(args.warmup, 1))
    v = 1.0 / (PASSES_PER_STEP * per_pass)
    sample = (f"{args.steps} timed samples of {t_sample:.1f} s: conv_in + down_blocks.0-2 of a batch-1 SDXL UNet pass of the fp32 "
              f"CPU oracle = {100 * frac:.1f}% of the pass FLOPs, extrapolated by FLOPs to {per_pass:.1f} s/pass; one step = "
              f"{PASSES_PER_STEP} passes; blend/CFG (<0.1%) and the VAE colour guidance are NOT included (conservative)")
    print(json.dumps({
        "impl": "reference", "metric": "denoising steps/sec SDXL 1024^2 5-region", "value": v, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_product(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from rtti_b200 import ops
    from rtti_b200.region_diffusion_sdxl import RegionDiffusionXL
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    model = RegionDiffusionXL.from_synthetic(seed=0, device=dev, with_vae=True)
    wl = synth_workload(dev)
    model.masks = [m.to(dev) for m in wl["masks"]]
    model.scheduler.set_timesteps(NUM_INFERENCE_STEPS)
    timesteps = model.scheduler.timesteps
    time_ids = torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]], device=dev)

    def fresh_state(src):
        tfd = dict(src["tfd"])
        tfd["color_obj_atten"] = [m.to(dev, non_blocking=True) for m in src["tfd"]["color_obj_atten"]]
        tfd["color_obj_atten_all"] = src["tfd"]["color_obj_atten_all"].to(dev, non_blocking=True)
        tfd["target_RGB"] = [r.to(dev, non_blocking=True) for r in src["tfd"]["target_RGB"]]
        lat = src["latents"].to(dev, torch.float16, non_blocking=True) * model.scheduler.init_noise_sigma
        model.masks = [m.to(dev, non_blocking=True) for m in src["masks"]]
        return model.prepare_rich_text(src["ctx"].to(dev, torch.float16, non_blocking=True),
                                       src["pooled"].to(dev, torch.float16, non_blocking=True), time_ids, lat, timesteps,
                                       GUIDANCE, True, 0.5, 0.5, tfd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------- device-resident timing
    st = fresh_state(wl)
    with torch.no_grad():
        for i in range(args.warmup):
            model.rich_text_step(st, i % NUM_INFERENCE_STEPS)
        barrier()
        clocks = ClockSampler(local_rank) if rank == 0 else None
        launches0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.cuda.nvtx.range_push("timed")
        for i in range(args.warmup, args.warmup + args.steps):
            model.rich_text_step(st, i % NUM_INFERENCE_STEPS)
        torch.cuda.nvtx.range_pop()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ops.LAUNCHES - launches0
        clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    assert bool(torch.isfinite(st.latents.float()).all()), "non-finite latents"

    # ------------------------------------------------------------- end to end: host buffers every step
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in wl.items()}
    pinned["masks"] = [m.pin_memory() for m in wl["masks"]]
    pinned["tfd"] = dict(wl["tfd"])
    pinned["tfd"]["color_obj_atten"] = [m.pin_memory() for m in wl["tfd"]["color_obj_atten"]]
    pinned["tfd"]["color_obj_atten_all"] = wl["tfd"]["color_obj_atten_all"].pin_memory()
    h2d = sum(x.numel() * x.element_size() for x in [pinned["ctx"], pinned["pooled"], pinned["latents"], *pinned["masks"],
                                                     *pinned["tfd"]["color_obj_atten"], pinned["tfd"]["color_obj_atten_all"]])
    host_lat = torch.empty(1, 4, 128, 128, dtype=torch.float16).pin_memory()
    with torch.no_grad():
        barrier()
        t0 = time.perf_counter()
        for i in range(args.warmup, args.warmup + args.steps):
            s2 = fresh_state(pinned)                 # H2D of this step's inputs from pinned host memory
            s2.kv_caches = st.kv_caches              # prompt K/V projections and the captured UNet graphs are
            s2.graphs = st.graphs                    # per-prompt state, kept across steps
            model.rich_text_step(s2, i % NUM_INFERENCE_STEPS)   # the public step call
            host_lat.copy_(s2.latents, non_blocking=False)      # D2H of the step result
            loss = float(model.last_step_stats["color_loss"].item())
        barrier()
        e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    d2h = host_lat.numel() * 2 + 4

    # ------------------------------------------------------------- roofline of the dominant rtti kernel (CUDA events)
    # (every rank executes the profiling steps: on >1 GPU each step contains the cross-rank exchange)
    roof = cross = None
    ops.PROFILE = []
    model.profile_events = {}
    graphs_on, model.use_cuda_graphs = model.use_cuda_graphs, False   # eager so every launch carries its events
    with torch.no_grad():
        model.rich_text_step(st, (args.warmup + args.steps) % NUM_INFERENCE_STEPS)
    barrier()
    prof, ops.PROFILE = ops.PROFILE, None
    breakdown = {k: a.elapsed_time(b) for k, (a, b) in model.profile_events.items()}
    breakdown["note"] = "eager (no CUDA graph) profiling step"
    model.use_cuda_graphs = graphs_on
    model.profile_events = {}
    with torch.no_grad():
        model.rich_text_step(st, (args.warmup + args.steps + 1) % NUM_INFERENCE_STEPS)
    barrier()
    breakdown_graph = {k: a.elapsed_time(b) for k, (a, b) in model.profile_events.items()}
    model.profile_events = None
    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        agg = {}
        for ev0, ev1, kind, flops, nbytes, shape in prof:
            a = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
            a[0] += ev0.elapsed_time(ev1) * 1e-3; a[1] += flops; a[2] += nbytes; a[3] += 1
        s = agg.get("self")
        c = agg.get("cross")
        if s:
            ach = s[1] / s[0] / 1e12
            roof = {"kernel": "attn_self_v3_kernel (self-attention, tcgen05/TMEM, head_dim 64)", "bound": "tensor", "achieved": ach,
                    "peak": tf_sust, "unit": "TFLOP/s", "frac": ach / tf_sust, "traffic": 68.73e6,
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum = 62.95 + 5.77 MB for one launch of the "
                                    "32x32-level shape (B8 h20 T1024 d64: 60 of the 70 launches of a step; algorithmic Q+K+V+O "
                                    "= 83.9 MB, part of O still in L2 at kernel end), ncu --set full capture "
                                    "profiles/r01_attn_self_v3_ncu_details.txt",
                    "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
                    "launches_timed": s[3], "ms_per_step_in_kernel": s[0] * 1e3}
        if c:
            gbs = c[2] / c[0] / 1e9
            cross = {"kernel": "attn_fwd_kernel<80,1> (cross-attention, 77 keys)", "bound": "hbm", "achieved": gbs,
                     "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "tensor_tflops": c[1] / c[0] / 1e12,
                     "launches_timed": c[3], "ms_per_step_in_kernel": c[0] * 1e3}

    if rank != 0:
        return
    steps_per_s = args.steps / (ms * 1e-3)
    line = {
        "metric": "denoising steps/sec SDXL 1024^2 5-region", "value": steps_per_s, "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": bench_config(),
        "execution": "the passes of a rank run as one batched, CUDA-graph-replayed UNet call",
        "parallelism": (f"UNet passes region-parallel x{world} (fused peer-memory exchange), colour guidance stripe-parallel x{world}"
                        if world > 1 else "single GPU"),
        "clocks": clk, "gpu_launches": launches,
        "e2e": {"value": args.steps / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "last_color_loss": loss},
        "roofline": roof, "roofline_cross_attention": cross, "breakdown_ms": breakdown_graph, "breakdown_eager_ms": breakdown,
    }
    if world == 1 and not args.no_cpu_baseline:
        per_pass, threads, t_sample, frac = cpu_oracle_pass_time(1, 1)
        v = 1.0 / (PASSES_PER_STEP * per_pass)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": threads, "kind": "port",
                                "sample": f"conv_in + down_blocks.0-2 of one batch-1 SDXL UNet pass of the fp32 CPU oracle "
                                          f"({t_sample:.1f} s = {100 * frac:.1f}% of the pass FLOPs), extrapolated by FLOPs to "
                                          f"{per_pass:.1f} s/pass x {PASSES_PER_STEP} passes/step; VAE colour guidance and blend "
                                          "not included (conservative)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rtti", choices=["rtti", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_product(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

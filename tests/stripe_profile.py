"""Per-kernel timeline of the stripe-parallel colour-guidance engine at SDXL size (torch.profiler / CUPTI, rank 0):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/stripe_profile.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cudnn.benchmark = os.environ.get("RTTI_CUDNN_BENCHMARK", "1") == "1"
    from rtti_b200.stripe_parallel import StripedDecoderFwdBwd
    from rtti_b200.vae import AutoencoderKLDecoder, VAEConfig
    vae = AutoencoderKLDecoder(VAEConfig.sdxl()).init_synthetic(seed=5).finalize("cuda")
    eng = StripedDecoderFwdBwd(vae, 128, 128, torch.device("cuda"))
    z = torch.randn(1, 4, 128, 128, device="cuda")
    wgt = torch.randn(1, 3, 1024, 1024, device="cuda")
    grad_fn = lambda img: torch.tanh(img) * wgt
    for _ in range(4):
        eng.backward(grad_fn(eng.forward(z)))
    torch.cuda.synchronize()
    dist.barrier()
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        eng.backward(grad_fn(eng.forward(z)))
    t_issue = (time.perf_counter() - t0) / 5
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"world={world}: {e0.elapsed_time(e1) / 5:.2f} ms per fwd+bwd on the GPU; CPU issue time {t_issue * 1e3:.2f} ms", flush=True)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        eng.backward(grad_fn(eng.forward(z)))
        torch.cuda.synchronize()
    eng.arena.check()
    # per-rank totals of the kernels that matter for the skew analysis (every rank profiles itself)
    tot = {}
    for e in prof.key_averages():
        t = getattr(e, "self_device_time_total", None)
        if t is None:
            t = e.self_cuda_time_total
        for key in ("gn32_finalize_peer_kernel<0>", "gn32_finalize_peer_kernel<1>", "halo_exchange", "implicit_gemm", "conv3d_fprop",
                    "gn32_partial", "gn32_apply", "ncclDevKernel"):
            if key in e.key:
                tot[key] = tot.get(key, 0.0) + t / 1e3
    for r in range(world):
        if r == rank:
            print(f"rank {rank} (cudnn.benchmark={torch.backends.cudnn.benchmark}) kernel ms: " +
                  ", ".join(f"{k}={v:.2f}" for k, v in sorted(tot.items())), flush=True)
        dist.barrier()
    if rank == 0 and os.environ.get("RTTI_PROFILE_TABLE", "0") == "1":
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

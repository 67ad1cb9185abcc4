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
    torch.backends.cudnn.benchmark = True
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
    if rank == 0:
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

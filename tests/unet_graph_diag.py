"""Diagnostic: SDXL UNet batched pass — eager wall time vs CPU launch time vs CUDA-graph replay."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtti_b200.unet import CrossKVCache, RegionControl, UNet2DConditionModel, UNetConfig  # noqa: E402


def main():
    dev = "cuda"
    with torch.device(dev):
        unet = UNet2DConditionModel(UNetConfig.sdxl())
    unet.finalize(dev).init_synthetic(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, 4, 128, 128, device=dev, generator=g).half()
    ctx = torch.randn(B, 77, 2048, device=dev, generator=g).half()
    added = {"text_embeds": torch.randn(B, 1280, device=dev, generator=g).half(),
             "time_ids": torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]], device=dev)}
    t_dev = torch.full((1,), 981.0, device=dev)
    kv = CrossKVCache()
    src = [0, 1, 2, 3] + [3] * (B - 4) if B > 4 else None
    idx = torch.as_tensor(src, device=dev) if src else None

    def run():
        ctrl = RegionControl(kv_cache=kv, qk_src=src, feature_src=src, feature_idx=idx)
        return unet(x, t_dev, ctx, added, ctrl)["sample"]

    with torch.no_grad():
        for _ in range(2):
            y = run()
        torch.cuda.synchronize()
        for name in ("eager",):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(3):
                y = run()
            e1.record(); t_cpu = time.perf_counter() - t0
            torch.cuda.synchronize()
            print(f"{name}: gpu wall {e0.elapsed_time(e1) / 3:.1f} ms/pass, cpu launch {t_cpu / 3 * 1e3:.1f} ms/pass", flush=True)
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(gr):
            yg = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gr.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"graph replay: {e0.elapsed_time(e1) / 5:.1f} ms/pass; max|graph-eager| = {(yg.float() - y.float()).abs().max().item():.3e}")
        print("mem GB", torch.cuda.max_memory_allocated() / 2 ** 30)


if __name__ == "__main__":
    main()

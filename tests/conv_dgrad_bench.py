"""Micro-benchmark (GPU): data gradient of the VAE decoder's 3x3 convolutions, cuDNN `convolution_backward`
(output_mask = input only) vs the same gradient evaluated as a forward convolution with the flipped, transposed
filter (what stripe_parallel.py does). fp32 channels-last, TF32 tensor cores.  python tests/conv_dgrad_bench.py"""
import json

import torch
import torch.nn.functional as F


def bench(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    torch.backends.cudnn.benchmark = True
    bwd = torch.ops.aten.convolution_backward
    for (hw, cin, cout) in ((1024, 128, 128), (1024, 256, 128), (1024, 256, 256), (512, 256, 256), (512, 512, 256),
                            (512, 512, 512), (256, 512, 512), (128, 512, 512), (1024, 128, 3)):
        w = torch.randn(cout, cin, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last)
        wf = w.flip(2, 3).transpose(0, 1).contiguous(memory_format=torch.channels_last)
        g = torch.randn(1, cout, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
        x = torch.empty(1, cin, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
        t_bwd = bench(lambda: bwd(g, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]))
        t_flip = bench(lambda: F.conv2d(g, wf, None, 1, 1))
        t_fwd = bench(lambda: F.conv2d(x, w, None, 1, 1))
        a = bwd(g, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False])[0]
        b = F.conv2d(g, wf, None, 1, 1)
        fl = 2 * hw * hw * cin * cout * 9
        print(json.dumps({"hw": hw, "cin": cin, "cout": cout, "dgrad_ms": round(t_bwd, 3), "flipped_fprop_ms": round(t_flip, 3),
                          "fprop_ms": round(t_fwd, 3), "dgrad_TFLOPs": round(fl / t_bwd / 1e9, 1),
                          "flipped_TFLOPs": round(fl / t_flip / 1e9, 1),
                          "rel_diff": float((a - b).abs().max() / a.abs().max())}), flush=True)
        del w, wf, g, x, a, b
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

"""Debug helper (GPU): per-layer row sums of the SD1.5-shaped token-map capture vs the golden, for A/B runs with
RTTI_ATTN_POLY=0/4 (see tests/test_parity_gpu.py::test_sd_loops_vs_reference_golden)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import synth  # noqa: E402
from tests.test_parity_gpu import _product_unet  # noqa: E402


def main():
    from oracle import unet_oracle as uo
    from rtti_b200.region_diffusion import RegionDiffusion
    g = np.load(os.path.join(ROOT, "tests", "golden", "sd_loops.npz"))
    cfg = uo.tiny_sd_config()
    S = 64
    model = RegionDiffusion(device="cuda", unet=_product_unet(cfg, 1), vae=synth.TinyVAE("cuda"))
    inp = synth.synth_inputs(cfg.cross_attention_dim, 0, 3, S, 21)
    ctx = inp["ctx"].cuda()
    model.register_tokenmap_hooks()
    model.produce_attn_maps(None, None, height=S * 8, width=S * 8, num_inference_steps=12, guidance_scale=8.5,
                            latents=inp["latents"].clone(), text_embeddings=torch.cat([ctx[:1], ctx[-1:]]), decode=False)
    names = sorted(model.selfattn_maps)
    for k, ref in zip(names, g["plain_self_rowsum"]):
        m = model.selfattn_maps[k][0]
        rs = m.sum(-1)
        bad = (rs - ref).abs() > 5e-3
        print(f"POLY={os.environ.get('RTTI_ATTN_POLY', 'default')} {k}: T={m.shape[0]} row0 sum {float(rs[0]):.4f} (golden {float(ref):.4f}); "
              f"rows off: {int(bad.sum())}/{m.shape[0]}; first bad rows {bad.nonzero().flatten()[:8].tolist()}; max rowsum {float(rs.max()):.3f}")


if __name__ == "__main__":
    main()

"""One launch of each hot rtti kernel at the SDXL batch-8 shapes, for `ncu --set full` (tools/gpu/*.sh)."""
import sys

import torch

sys.path.insert(0, ".")
from rtti_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g).half()
B = 8
for H, T in ((20, 1024), (10, 4096)):
    C = H * 64
    qkv = rn(B, T, 3 * C)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    o = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
    kc, vc, qc = rn(B, 77, C), rn(B, 77, C), rn(B, T, C)
    for _ in range(2):
        ops.attention(q, k, v, H, out=o)                                            # attn_self_kernel<1, 4>
        ops.attention(q, k, v, H, out=o, qk_src=[0, 1, 2, 3, 3, 3, 3, 3])            # attn_self_kernel<6, 0> + <1, 4>
        ops.attention(qc, kc, vc, H, out=o)                                         # attn_cross_kernel
for rows, C in ((32768, 640), (8192, 1280)):
    x = rn(rows // 8, 8, C).reshape(8, rows // 8, C)
    w = rn(8 * C, C) / 25
    b = rn(8 * C)
    for _ in range(2):
        ops.ff_geglu(x, w, b)
    a = rn(8, rows // 8, C)
    ga, be = rn(C), rn(C)
    for _ in range(2):
        ops.add_bias_layernorm(a, x.clone(), b[:C].contiguous(), ga, be, 1e-5)
torch.cuda.synchronize()

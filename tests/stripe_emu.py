"""Torch emulations of the stripe-parallel kernels' C-ABI contracts (include/rtti_b200.h: rtti_gn32_silu_fwd/bwd,
rtti_gn32_silu_*_striped, rtti_add_bias_f32) for the CPU tests of rtti_b200/stripe_parallel.py. Test infrastructure
only: the product path never imports this module."""
import torch
import torch.nn.functional as F

from rtti_b200.vae import AutoencoderKLDecoder, VAEConfig


def make_vae():
    cfg = VAEConfig(block_out_channels=(8, 8, 16, 16), norm_num_groups=4)
    vae = AutoencoderKLDecoder(cfg).init_synthetic(seed=3).float().eval()
    g = torch.Generator().manual_seed(17)
    for p in vae.parameters():   # non-trivial biases / affine parameters
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn(p.shape, generator=g))
    vae.requires_grad_(False)
    return vae


def autograd_reference(vae, zs, grad_fn):
    want = []
    for z in zs:
        zz = z.clone().requires_grad_(True)
        with torch.enable_grad():
            img = vae.decode_tensor(zz)
        img.backward(grad_fn(img.detach()))
        want.append((img.detach(), zz.grad))
    return want


def _group_sums(v, groups):   # v [hw, C] -> [groups]
    return v.view(v.shape[0], groups, -1).sum(dim=(0, 2))


def fake_ops(reduce_over_ranks):
    """name -> emulation. reduce_over_ranks(key, tensor) returns the sum of `tensor` over the ranks (the peer
    reduction of gn32_finalize_peer_kernel); every rank must call it in the same order."""
    def gn_stats(x, cb, groups, n, eps, red):
        xs = x[0] + (cb if cb is not None else 0)
        s0, s1 = red(_group_sums(xs, groups)), red(_group_sums(xs * xs, groups))
        mean = s0 / n
        rstd = torch.rsqrt((s1 / n - mean * mean).clamp_min(0) + eps)
        return xs, torch.stack([mean, rstd], 1)[None]

    def gn_fwd(x, gamma, beta, groups, eps, silu, n, red, chan_bias, out):
        xs, stats = gn_stats(x, chan_bias, groups, n, eps, red)
        cpg = x.shape[2] // groups
        y = (xs - stats[0, :, 0].repeat_interleave(cpg)) * stats[0, :, 1].repeat_interleave(cpg) * gamma + beta
        y = F.silu(y) if silu else y
        if out is None:
            out = torch.empty_like(x)
        out.view_as(x).copy_(y[None])
        return out.view_as(x), stats

    def gn_bwd(x, dz, gamma, beta, stats, groups, silu, n, red, chan_bias, out):
        cpg = x.shape[2] // groups
        xs = x[0] + (chan_bias if chan_bias is not None else 0)
        mu, rs = stats[0, :, 0].repeat_interleave(cpg), stats[0, :, 1].repeat_interleave(cpg)
        xh = (xs - mu) * rs
        dy = dz.reshape(xs.shape)
        if silu:
            y = xh * gamma + beta
            sg = torch.sigmoid(y)
            dy = dy * sg * (1 + y * (1 - sg))
        t = dy * gamma
        c1 = (red(_group_sums(t, groups)) / n).repeat_interleave(cpg)
        c2 = (red(_group_sums(t * xh, groups)) / n).repeat_interleave(cpg)
        dx = rs * (t - c1 - xh * c2)
        if out is None:
            out = torch.empty_like(x)
        out.view_as(x).copy_(dx[None])
        return out.view_as(x)

    ident = lambda v: v

    def add_bias(a, b, bias=None, out=None):
        r = a + b + (bias if bias is not None else 0)
        return r if out is None else out.copy_(r)

    return {
        "gn32_silu_fwd": lambda x, g, b, groups, eps, silu, chan_bias=None:
            gn_fwd(x, g, b, groups, eps, silu, x.shape[1] * x.shape[2] // groups, ident, chan_bias, None),
        "gn32_silu_bwd": lambda x, dz, g, b, stats, groups, silu, chan_bias=None:
            gn_bwd(x, dz, g, b, stats, groups, silu, x.shape[1] * x.shape[2] // groups, ident, chan_bias, None),
        "gn32_silu_fwd_striped": lambda x, g, b, groups, eps, silu, hw_total, peers, seq, chan_bias=None, out=None:
            gn_fwd(x, g, b, groups, eps, silu, hw_total * x.shape[2] // groups,
                   lambda v: reduce_over_ranks(("gn", seq), v), chan_bias, out),
        "gn32_silu_bwd_striped": lambda x, dz, g, b, stats, groups, silu, hw_total, peers, seq, chan_bias=None, out=None:
            gn_bwd(x, dz, g, b, stats, groups, silu, hw_total * x.shape[2] // groups,
                   lambda v: reduce_over_ranks(("gn", seq), v), chan_bias, out),
        "add_bias_f32": add_bias,
    }


class FakeArenaBase:
    """Interface of stripe_parallel.StripeArena on CPU tensors; subclasses implement exchange()."""

    def __init__(self, world, rank, pad_bytes, group=None):
        self.world, self.rank, self.group = world, rank, group
        self.pad_bytes = pad_bytes
        self.halves = [torch.full((pad_bytes // 4,), float("nan")) for _ in range(2)]   # NaN: an unwritten halo shows up
        self.gn_seq = self.halo_seq = 0

    def next_gn_seq(self):
        self.gn_seq += 1
        return self.gn_seq

    def pad(self, rows, W, C):
        self.halo_seq += 1
        n = (rows + 2) * W * C
        assert n * 4 <= self.pad_bytes
        return self.halves[self.halo_seq & 1][:n].view(rows + 2, W, C), self.halo_seq

    def release(self, seq):
        assert seq == self.halo_seq
        self.halo_seq -= 1

    def check(self):
        pass


def assert_matches(results, want, rank0_results=None):
    for k, ((img, g), (img_w, g_w)) in enumerate(zip(results, want)):
        assert torch.allclose(img, img_w, rtol=1e-4, atol=1e-4 * float(img_w.abs().max()))
        assert torch.allclose(g, g_w, rtol=1e-3, atol=1e-4 * float(g_w.abs().max()))
        if rank0_results is not None:
            assert torch.equal(g, rank0_results[k][1])   # broadcast: identical on all ranks

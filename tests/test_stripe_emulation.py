"""Host logic of the stripe-parallel colour-guidance engine (rtti_b200/stripe_parallel.py) on CPU, world 2 and 4.

The CUDA kernels and the symmetric-memory arena are replaced by torch emulations that follow the C-ABI contracts of
include/rtti_b200.h (rtti_gn32_silu_*_striped, rtti_halo_exchange, rtti_add_bias_f32), the ranks are threads, and
the collectives are barriers. What is tested is the orchestration the GPU cannot check cheaply: stripe bookkeeping
and tape order, halo rows, the flipped-filter data gradient, global GroupNorm statistics, pad re-use by parity —
against plain autograd through the same decoder. The kernels themselves are covered by tests/multigpu_check.py."""
import threading

import pytest
import torch
import torch.nn.functional as F

from rtti_b200 import ops, stripe_parallel
from rtti_b200.vae import AutoencoderKLDecoder, VAEConfig


class _World:
    def __init__(self, n):
        self.n = n
        self.barrier = threading.Barrier(n)
        self.arenas = [None] * n
        self.slots = {}
        self.lock = threading.Lock()
        self.local = threading.local()


class _FakeDist:
    """all_gather_into_tensor / broadcast over threads."""

    def __init__(self, w):
        self.w = w

    def get_world_size(self, group=None):
        return self.w.n

    def get_rank(self, group=None):
        return self.w.local.rank

    def get_global_rank(self, group, r):
        return r

    def all_gather_into_tensor(self, out, inp, group=None):
        w, r = self.w, self.w.local.rank
        w.slots[("ag", r)] = inp.clone()
        w.barrier.wait()
        out.copy_(torch.cat([w.slots[("ag", k)].reshape(-1) for k in range(w.n)]))
        w.barrier.wait()

    def reduce_scatter_tensor(self, out, inp, group=None):
        w, r = self.w, self.w.local.rank
        w.slots[("rs", r)] = inp.clone()
        w.barrier.wait()
        tot = sum(w.slots[("rs", k)] for k in range(w.n))
        out.copy_(tot.view(w.n, -1)[r])
        w.barrier.wait()

    def broadcast(self, t, src, group=None):
        w, r = self.w, self.w.local.rank
        if r == src:
            w.slots["bc"] = t.clone()
        w.barrier.wait()
        t.copy_(w.slots["bc"])
        w.barrier.wait()


class _FakeArena:
    """Same interface as stripe_parallel.StripeArena, CPU tensors, neighbours reached through the _World."""

    def __init__(self, w, rank, pad_bytes):
        self.w, self.world, self.rank, self.group = w, w.n, rank, None
        self.pad_bytes = pad_bytes
        self.halves = [torch.full((pad_bytes // 4,), float("nan")) for _ in range(2)]   # NaN: unwritten halo shows up
        self.gn_seq = self.halo_seq = 0
        w.arenas[rank] = self

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

    def exchange(self, pad, seq):
        assert seq == self.halo_seq
        w, r = self.w, self.rank
        rows = pad.shape[0] - 2
        w.barrier.wait()   # every rank has written its interior (the kernel's pushes only need the own interior)
        if r > 0:
            w.arenas[r - 1].halves[seq & 1][:pad.numel()].view_as(pad)[rows + 1].copy_(pad[1])
        else:
            pad[0].zero_()
        if r + 1 < w.n:
            w.arenas[r + 1].halves[seq & 1][:pad.numel()].view_as(pad)[0].copy_(pad[rows])
        else:
            pad[rows + 1].zero_()
        w.barrier.wait()   # "wait for the neighbours' flags"

    def check(self):
        pass


def _group_sums(v, groups):   # v [hw, C] -> [groups]
    return v.view(v.shape[0], groups, -1).sum(dim=(0, 2))


def _install_fake_ops(monkeypatch, w):
    def reduce_over_ranks(key, val):
        r = w.local.rank
        w.slots[(key, r)] = val
        w.barrier.wait()
        tot = sum(w.slots[(key, k)] for k in range(w.n))
        w.barrier.wait()
        return tot

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
    monkeypatch.setattr(ops, "gn32_silu_fwd", lambda x, g, b, groups, eps, silu, chan_bias=None:
                        gn_fwd(x, g, b, groups, eps, silu, x.shape[1] * x.shape[2] // groups, ident, chan_bias, None))
    monkeypatch.setattr(ops, "gn32_silu_bwd", lambda x, dz, g, b, stats, groups, silu, chan_bias=None:
                        gn_bwd(x, dz, g, b, stats, groups, silu, x.shape[1] * x.shape[2] // groups, ident, chan_bias, None))
    monkeypatch.setattr(ops, "gn32_silu_fwd_striped",
                        lambda x, g, b, groups, eps, silu, hw_total, peers, seq, chan_bias=None, out=None:
                        gn_fwd(x, g, b, groups, eps, silu, hw_total * x.shape[2] // groups,
                               lambda v: reduce_over_ranks(("gn", seq), v), chan_bias, out))
    monkeypatch.setattr(ops, "gn32_silu_bwd_striped",
                        lambda x, dz, g, b, stats, groups, silu, hw_total, peers, seq, chan_bias=None, out=None:
                        gn_bwd(x, dz, g, b, stats, groups, silu, hw_total * x.shape[2] // groups,
                               lambda v: reduce_over_ranks(("gn", seq), v), chan_bias, out))

    def add_bias(a, b, bias=None, out=None):
        r = a + b + (bias if bias is not None else 0)
        return r if out is None else out.copy_(r)
    monkeypatch.setattr(ops, "add_bias_f32", add_bias)


@pytest.mark.parametrize("world", [2, 4])
def test_striped_decoder_matches_autograd(monkeypatch, world):
    torch.manual_seed(0)
    torch.set_num_threads(1)
    cfg = VAEConfig(block_out_channels=(8, 8, 16, 16), norm_num_groups=4)
    vae = AutoencoderKLDecoder(cfg).init_synthetic(seed=3).float().eval()
    for p in vae.parameters():   # non-trivial biases / affine parameters
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    vae.requires_grad_(False)
    h = wd = 8
    w = _World(world)
    _install_fake_ops(monkeypatch, w)
    zs = [torch.randn(1, 4, h, wd) for _ in range(2)]
    wgt = torch.randn(1, 3, 8 * h, 8 * wd)
    grad_fn = lambda img: torch.tanh(img) * wgt

    want = []
    for z in zs:
        zz = z.clone().requires_grad_(True)
        with torch.enable_grad():
            img = vae.decode_tensor(zz)
        img.backward(grad_fn(img.detach()))
        want.append((img.detach(), zz.grad))

    pad_bytes = stripe_parallel.stripe_pad_elems(vae.decoder, h // world, wd) * 4
    out, errs = [None] * world, []

    def run(rank):
        try:
            w.local.rank = rank
            eng = stripe_parallel.StripedDecoderFwdBwd(vae, h, wd, "cpu", arena=_FakeArena(w, rank, pad_bytes), dist=_FakeDist(w))
            res = []
            for z in zs:   # two calls: pads / sequence numbers carry over
                img = eng.forward(z)
                res.append((img.clone(), eng.backward(grad_fn(img))))
            out[rank] = res
        except BaseException as e:   # noqa: BLE001 - re-raised in the main thread
            errs.append(e)
            w.barrier.abort()

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ths]
    [t.join(120) for t in ths]
    assert not errs, errs
    for rank in range(world):
        for k, ((img, g), (img_w, g_w)) in enumerate(zip(out[rank], want)):
            assert torch.allclose(img, img_w, rtol=1e-4, atol=1e-4 * float(img_w.abs().max()))
            assert torch.allclose(g, g_w, rtol=1e-3, atol=1e-4 * float(g_w.abs().max()))
            assert torch.equal(g, out[0][k][1])   # broadcast: identical on all ranks

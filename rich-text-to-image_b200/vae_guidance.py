"""Explicit forward + input-gradient backward of the VAE decoder for colour guidance.

The reference back-propagates `loss -> imgs -> vae.decode -> latents` with autograd through the third-party
fp32 AutoencoderKL (models/region_diffusion_sdxl.py:849-867, models/region_diffusion.py:151-168), including the
weight gradients it never uses. Here the decoder (vae.py, diffusers parameter names) is evaluated as an explicit
static sequence — no autograd graph, no autograd thread, CUDA-graph capturable:

  * channels-last fp32 activations [B, H*W, C] end to end;
  * GroupNorm(+SiLU) forward and backward in the sm_100a kernels of csrc/vae_kernels.cu
    (PyTorch's native GroupNorm round-trips channels-last tensors through NCHW copies);
  * 3x3 / 1x1 convolutions: cuDNN forward and `convolution_backward` with output_mask=(True, False, False)
    (data gradient only, TF32 tensor cores as in the reference's default PyTorch settings);
  * the single-head 16384-token mid-block attention materialises its 1 GB probability matrix once
    (B200: 180 GB) and reuses it for the five backward GEMMs instead of recomputing it.
"""
import math

import torch
import torch.nn.functional as F

from . import ops

_conv_bwd = torch.ops.aten.convolution_backward


class _Tape(list):
    pass


def _nchw(x, H, W):
    """[B, HW, C] contiguous -> logical NCHW view with channels_last strides (no copy)."""
    return x.view(x.shape[0], H, W, x.shape[2]).permute(0, 3, 1, 2)


def _cl(y):
    """NCHW (channels_last memory) -> [B, HW, C] contiguous view."""
    B, C, H, W = y.shape
    y = y.permute(0, 2, 3, 1)
    if not y.is_contiguous():
        y = y.contiguous()
    return y.view(B, H * W, C), H, W


def _conv_f(conv, x, H, W, with_bias=True):
    """with_bias=False: the caller folds conv.bias into the consumer (GroupNorm kernel / fused residual add) —
    PyTorch otherwise adds the bias of a channels-last fp32 convolution in a separate broadcast pass."""
    y = F.conv2d(_nchw(x, H, W), conv.weight, conv.bias if with_bias else None, conv.stride, conv.padding)
    return _cl(y)[0]


class DecoderFwdBwd:
    """decode(z) -> image, then backward(d image) -> d z, for vae.AutoencoderKLDecoder parameters."""

    def __init__(self, vae):
        self.vae = vae
        self.groups = vae.config.norm_num_groups
        self._dummies = {}

    # ------------------------------------------------------------------ pieces
    def _gn_f(self, norm, x, silu, tape, chan_bias=None):
        y, stats = ops.gn32_silu_fwd(x, norm.weight, norm.bias, self.groups, norm.eps, silu, chan_bias=chan_bias)
        tape.append(("gn", norm, x, stats, silu, chan_bias))
        return y

    def _gn_b(self, rec, g):
        _, norm, x, stats, silu, chan_bias = rec
        return ops.gn32_silu_bwd(x, g.contiguous(), norm.weight, norm.bias, stats, self.groups, silu, chan_bias=chan_bias)

    def _dummy(self, shape, dev):
        k = (tuple(shape), str(dev))
        if k not in self._dummies:
            self._dummies[k] = torch.empty(shape, dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        return self._dummies[k]

    def _conv_b(self, conv, g, cin, H, W):
        g4 = _nchw(g, H, W)
        gi, _, _ = _conv_bwd(g4, self._dummy((g.shape[0], cin, H, W), g.device), conv.weight, None, list(conv.stride),
                             list(conv.padding), [1, 1], False, [0, 0], 1, [True, False, False])
        return _cl(gi)[0]

    def _resnet_f(self, r, x, H, W, tape):
        h = self._gn_f(r.norm1, x, True, tape)
        h = _conv_f(r.conv1, h, H, W, with_bias=False)
        h = self._gn_f(r.norm2, h, True, tape, chan_bias=r.conv1.bias)   # conv1 bias folded into the norm
        h = _conv_f(r.conv2, h, H, W, with_bias=False)
        sc = _conv_f(r.conv_shortcut, x, H, W) if r.conv_shortcut is not None else x
        tape.append(("res", r, H, W))
        return ops.add_bias_f32(sc, h, r.conv2.bias)                      # residual + conv2 bias in one pass

    def _resnet_b(self, tape, g):
        _, r, H, W = tape.pop()
        cin, cout = r.conv1.in_channels, r.conv1.out_channels
        dh = self._conv_b(r.conv2, g, cout, H, W)
        dh = self._gn_b(tape.pop(), dh)
        dh = self._conv_b(r.conv1, dh, cin, H, W)
        dx = self._gn_b(tape.pop(), dh)
        if r.conv_shortcut is not None:
            return ops.add_bias_f32(dx, self._conv_b(r.conv_shortcut, g, cin, H, W))
        return ops.add_bias_f32(dx, g.contiguous())

    def _attn_f(self, a, x, tape):
        """Single-head self-attention over all tokens (vae._MidAttention), probabilities materialised."""
        B, T, C = x.shape
        hn = self._gn_f(a.group_norm, x, False, tape)
        q = F.linear(hn, a.to_q.weight, a.to_q.bias)
        k = F.linear(hn, a.to_k.weight, a.to_k.bias)
        v = F.linear(hn, a.to_v.weight, a.to_v.bias)
        scale = 1.0 / math.sqrt(C)
        # the softmax scale is applied to q ([T, C], 33 MB) instead of the [T, T] scores (1 GB at 1024^2: a 2 GB pass)
        p = torch.softmax(torch.bmm(q * scale, k.transpose(1, 2)), dim=-1)
        o = torch.bmm(p, v)
        tape.append(("attn", a, hn, q, k, v, p, scale))
        return x + F.linear(o, a.to_out[0].weight, a.to_out[0].bias)

    def _attn_b(self, tape, g):
        _, a, hn, q, k, v, p, scale = tape.pop()
        do = g @ a.to_out[0].weight                       # [B,T,C]
        dv = torch.bmm(p.transpose(1, 2), do)
        dp = torch.bmm(do, v.transpose(1, 2))
        ds = torch._softmax_backward_data(dp, p, -1, p.dtype)   # d/d(scaled scores); the scale goes onto the small operands
        dq = torch.bmm(ds, k * scale)
        dk = torch.bmm(ds.transpose(1, 2), q * scale)
        dhn = dq @ a.to_q.weight + dk @ a.to_k.weight + dv @ a.to_v.weight
        return g + self._gn_b(tape.pop(), dhn)

    # ------------------------------------------------------------------ whole decoder
    def forward(self, z):
        """z [B, 4, h, w] fp32 -> image [B, 3, 8h, 8w] fp32 (NCHW view of channels-last memory); keeps the tape."""
        vae, d = self.vae, self.vae.decoder
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            tape = _Tape()
            B, _, H, W = z.shape
            x = z.permute(0, 2, 3, 1).contiguous().view(B, H * W, -1)
            x = _conv_f(vae.post_quant_conv, x, H, W)
            x = _conv_f(d.conv_in, x, H, W)
            x = self._resnet_f(d.mid_block.resnets[0], x, H, W, tape)
            x = self._attn_f(d.mid_block.attentions[0], x, tape)
            x = self._resnet_f(d.mid_block.resnets[1], x, H, W, tape)
            for blk in d.up_blocks:
                for r in blk.resnets:
                    x = self._resnet_f(r, x, H, W, tape)
                if blk.upsamplers is not None:
                    C = x.shape[2]
                    x = x.view(B, H, 1, W, 1, C).expand(B, H, 2, W, 2, C).reshape(B, 4 * H * W, C)
                    H, W = 2 * H, 2 * W
                    x = _conv_f(blk.upsamplers[0].conv, x, H, W)
                    tape.append(("up", blk.upsamplers[0].conv, H, W, C))
            x = self._gn_f(d.conv_norm_out, x, True, tape)
            y = _conv_f(d.conv_out, x, H, W)
            tape.append(("out", H, W))
            self.tape = tape
            return _nchw(y, H, W)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def backward(self, grad_image):
        """grad_image [B, 3, H, W] -> d loss / d z [B, 4, h, w] fp32."""
        vae, d = self.vae, self.vae.decoder
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            tape = self.tape
            _, H, W = tape.pop()
            B = grad_image.shape[0]
            g = grad_image.permute(0, 2, 3, 1).contiguous().view(B, H * W, -1)
            g = self._conv_b(d.conv_out, g, d.conv_out.in_channels, H, W)
            g = self._gn_b(tape.pop(), g)
            for blk in reversed(d.up_blocks):
                if blk.upsamplers is not None:
                    _, conv, H, W, C = tape.pop()
                    g = self._conv_b(conv, g, C, H, W)
                    H, W = H // 2, W // 2
                    g = g.view(B, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(B, H * W, C)   # adjoint of nearest x2
                for _ in blk.resnets:
                    g = self._resnet_b(tape, g)
            g = self._resnet_b(tape, g)
            g = self._attn_b(tape, g)
            g = self._resnet_b(tape, g)
            g = self._conv_b(d.conv_in, g, d.conv_in.in_channels, H, W)
            g = self._conv_b(vae.post_quant_conv, g, vae.post_quant_conv.in_channels, H, W)
            self.tape = None
            return g.view(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev


def default_engine(vae):
    """The single-GPU explicit forward/backward engine of `vae` (created once, kept on the module)."""
    eng = getattr(vae, "_fwd_bwd", None)
    if eng is None:
        eng = vae._fwd_bwd = DecoderFwdBwd(vae)
    return eng


def image_and_latent_grad(vae, z, grad_fn, engine=None):
    """image = decode(z); grad_image = grad_fn(image.detach()); returns d/dz of <grad_image, decode(z)>.
    Uses the explicit forward/backward for vae.AutoencoderKLDecoder and autograd for any other decoder object
    exposing decode_tensor() (e.g. the tiny stand-in of the parity fixtures). `engine`: a
    stripe_parallel.StripedDecoderFwdBwd to run the decoder split by rows over the ranks."""
    from .vae import AutoencoderKLDecoder
    if isinstance(vae, AutoencoderKLDecoder):
        eng = engine if engine is not None else default_engine(vae)
        img = eng.forward(z)
        return eng.backward(grad_fn(img))
    z = z.detach().requires_grad_(True)
    with torch.enable_grad():
        img = vae.decode_tensor(z)
    img.backward(grad_fn(img.detach()))
    return z.grad


class GuidanceGraph:
    """decode -> colour loss forward/backward -> decoder backward as ONE replayed CUDA graph.

    The evaluation is a static sequence of ~600 launches (the stripe-parallel engine adds one halo kernel per convolution
    and three NCCL collectives); on 4-8 GPUs each rank's share of the GPU work shrinks to 7-15 ms while the CPU issue time
    stays ~15 ms, i.e. the phase is launch-bound unless it is replayed. Protocol per (engine, shapes): the first call runs
    eagerly (cuDNN autotune, arena views, flipped filters), the second captures and replays, later calls replay. Inputs
    are copied into static buffers; the returned gradient and loss are static tensors overwritten by the next call.
    Every rank of a stripe group must make the same sequence of calls (they do: the guidance is replicated control flow).
    """

    def __init__(self, engine):
        self.engine = engine
        self.calls = 0
        self.graph = None
        self.launches = 0

    def _run(self, z, masks, tgt):
        img = self.engine.forward(z)
        loss, g = ops.color_loss_fwd_bwd(img[0].contiguous(), masks, tgt)
        return loss, self.engine.backward(g[None])

    def __call__(self, z, masks, tgt):
        self.calls += 1
        if self.calls == 1:
            return self._run(z, masks, tgt)
        if self.graph is None:
            self.z, self.masks, self.tgt = z.clone(), masks.clone(), tgt.clone()
            torch.cuda.synchronize(z.device)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self.loss, self.grad = self._run(self.z, self.masks, self.tgt)
            self.launches = ops.LAUNCHES - n0
            self.graph = graph
        else:
            self.z.copy_(z); self.masks.copy_(masks); self.tgt.copy_(tgt)
        self.graph.replay()
        ops._count(self.launches)
        return self.loss, self.grad

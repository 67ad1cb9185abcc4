"""B200RegionAttnProcessor — plug-in for the REFERENCE's own `Attention` modules.

Contract of the reference (models/attention_processor.py:1114-1123, 1183; installed with
`unet.set_attn_processor(...)`, models/unet_2d_condition.py:594-626):

    processor(attn, hidden_states, real_attn_probs=None, attn_weights=None, encoder_hidden_states=None,
              attention_mask=None, temb=None) -> (hidden_states, [attention_probs_avg, attention_probs])

This processor keeps that contract while never materialising the probability tensor:
  * `attention_probs` (out[1][1]) is a `LazyAttentionProbs` handle carrying the call's Q and K; the reference's
    store hook keeps it (`out[1][1].detach()`, region_diffusion_sdxl.py:1082) and its replacement hook hands it
    back positionally as `real_attn_probs` (:1028) — injection then runs the fused kernel on (Q_ref, K_ref, V_new),
    which is exactly `P_ref @ V_new` (attention_processor.py:1160-1163);
  * `attention_probs_avg` (out[1][0]) is produced on request (`return_probs_avg=True`, what the token-map hook
    reads, :980-992) by the capture path of the kernels, as a device tensor;
  * `attn_weights={'word_pos','font_size'}` is the font-size path of attention_processor.py:387-399.
The whole-sampler path (unet.py) does not go through this class; it exists for users who keep the reference UNet.
"""
import torch

from . import ops


class LazyAttentionProbs:
    """Stands in for the [B*h, T, K] probability tensor: shape/dtype/detach() like a tensor, data = (q, k)."""

    def __init__(self, q, k, heads, scale):
        self.q, self.k, self.heads, self.scale = q, k, heads, scale
        self.shape = torch.Size((q.shape[0] * heads, q.shape[1], k.shape[1]))
        self.dtype, self.device = q.dtype, q.device

    def detach(self):
        return self

    def materialize(self):
        raise RuntimeError("rtti_b200 never materialises attention probabilities; use probs_avg or inject the handle")


class B200RegionAttnProcessor:
    def __init__(self, return_probs_avg=False):
        self.return_probs_avg = return_probs_avg

    def __call__(self, attn, hidden_states, real_attn_probs=None, attn_weights=None, encoder_hidden_states=None,
                 attention_mask=None, temb=None):
        if attention_mask is not None or getattr(attn, "spatial_norm", None) is not None or \
                getattr(attn, "group_norm", None) is not None or hidden_states.ndim != 3:
            raise NotImplementedError("B200RegionAttnProcessor covers the SD1.5/SDXL transformer attention sites only")
        residual = hidden_states
        hs = hidden_states.to(torch.float16)
        enc = hs if encoder_hidden_states is None else encoder_hidden_states.to(torch.float16)
        heads = attn.heads
        lin = torch.nn.functional.linear
        v = lin(enc, attn.to_v.weight.half())
        B, T, C = hs.shape
        if real_attn_probs is not None:
            if not isinstance(real_attn_probs, LazyAttentionProbs):
                raise TypeError("real_attn_probs must be the LazyAttentionProbs handle returned by this processor")
            q, k = real_attn_probs.q, real_attn_probs.k
        else:
            q = lin(hs, attn.to_q.weight.half())
            k = lin(enc, attn.to_k.weight.half())
        kw = {}
        if attn_weights is not None:
            assert k.shape[1] == 77
            kw = dict(word_pos=attn_weights["word_pos"].to(hs.device, torch.int32),
                      font_size=attn_weights["font_size"].to(hs.device, torch.float32), fs_batch_mask=(1 << B) - 1)
        probs_avg = None
        nk = k.shape[1]
        if self.return_probs_avg and nk <= 80:
            acc = torch.zeros(B, T, nk, dtype=torch.float32, device=hs.device)
            o = ops.attention(q, k, v, heads, scale=attn.scale, pbar_accum=acc, cap_slot=list(range(B)), **kw)
            probs_avg = acc.to(hidden_states.dtype)
        elif self.return_probs_avg:
            lse = torch.empty(B, heads, T, dtype=torch.float32, device=hs.device)
            o = ops.attention(q, k, v, heads, scale=attn.scale, lse=lse)
            acc = torch.zeros(B, T, nk, dtype=torch.float32, device=hs.device)
            for b in range(B):
                ops.attn_probs_mean_accum(q[b], k[b], lse[b], acc[b], heads, scale=attn.scale)
            probs_avg = acc.to(hidden_states.dtype)
        else:
            o = ops.attention(q, k, v, heads, scale=attn.scale, **kw)
        out = lin(o, attn.to_out[0].weight.half(), attn.to_out[0].bias.half() if attn.to_out[0].bias is not None else None)
        out = out.to(hidden_states.dtype)
        if getattr(attn, "residual_connection", False):
            out = out + residual
        out = out / getattr(attn, "rescale_output_factor", 1.0)
        return out, [probs_avg, LazyAttentionProbs(q, k, heads, attn.scale)]

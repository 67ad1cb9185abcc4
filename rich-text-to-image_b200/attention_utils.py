"""Token maps: layer allow-lists and get_token_maps (drop-in for utils/attention_utils.py:12-67, 233-341).

The captured maps arrive as fp32 DEVICE tensors (unet.TokenMapAccumulator) instead of fp16 CPU tensors:
averaging and the bicubic resizes run on the GPU; the spectral clustering stays the reference's own
scikit-learn call on the host (utils/attention_utils.py:262-265) so the segment labels are the ones the
reference would produce for the same affinity (SURVEY §8(f).3 lists a GPU eigensolver as "next").
The debug JPEG / matplotlib dumps of the reference are not reproduced.
"""
import random

import numpy as np
import torch

SelfAttentionLayers = [f"{b}.transformer_blocks.0.attn1" for b in (
    "down_blocks.0.attentions.0", "down_blocks.0.attentions.1", "down_blocks.1.attentions.0",
    "down_blocks.1.attentions.1", "down_blocks.2.attentions.0", "down_blocks.2.attentions.1",
    "mid_block.attentions.0", "up_blocks.1.attentions.0", "up_blocks.1.attentions.1", "up_blocks.1.attentions.2",
    "up_blocks.2.attentions.0", "up_blocks.2.attentions.1", "up_blocks.2.attentions.2", "up_blocks.3.attentions.0",
    "up_blocks.3.attentions.1", "up_blocks.3.attentions.2")]

CrossAttentionLayers = [f"{b}.transformer_blocks.0.attn2" for b in (
    "down_blocks.1.attentions.0", "down_blocks.2.attentions.0", "down_blocks.2.attentions.1",
    "mid_block.attentions.0", "up_blocks.1.attentions.0", "up_blocks.1.attentions.1", "up_blocks.1.attentions.2",
    "up_blocks.2.attentions.1")]

CrossAttentionLayers_XL = (
    [f"down_blocks.2.attentions.1.transformer_blocks.{i}.attn2" for i in (3, 4)]
    + [f"mid_block.attentions.0.transformer_blocks.{i}.attn2" for i in (0, 1, 2, 3)]
    + [f"up_blocks.0.attentions.0.transformer_blocks.{i}.attn2" for i in (1, 2, 3, 4, 5, 6, 7)]
    + ["up_blocks.1.attentions.0.transformer_blocks.0.attn2"])


def seed_everything(seed):
    """utils/richtext_utils.py:22-27."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _resize(x, size):
    return torch.nn.functional.interpolate(x, size, mode="bicubic", antialias=True)


def self_affinity(selfattn_maps, resolution=32):
    """utils/attention_utils.py:241-256: layers whose side is `resolution`, averaged -> [res^2, res^2] (device)."""
    keep = []
    for m in selfattn_maps.values():
        r = int(round(float(np.sqrt(m.shape[1]))))
        if r != resolution:
            continue
        a = m.reshape(1, r, r, r * r).permute(3, 0, 1, 2).float()
        a = _resize(a, (resolution, resolution))  # identity-sized resize, kept because antialiased bicubic is not a no-op
        keep.append(a.permute(1, 2, 3, 0).reshape(1, resolution ** 2, r * r))
    return torch.cat(keep).mean(0)


def cross_maps_mean(crossattn_maps, resolution=32):
    """utils/attention_utils.py:281-291 -> [res, res, 77] (device)."""
    outs = []
    for m in crossattn_maps.values():
        r = int(round(float(np.sqrt(m.shape[1]))))
        a = m.reshape(1, r, r, -1).permute(0, 3, 1, 2).float()
        outs.append(_resize(a, (resolution, resolution)).permute(0, 2, 3, 1))
    return torch.cat(outs).mean(0)


_PALETTE = np.array([[68, 1, 84], [59, 82, 139], [33, 145, 140], [94, 201, 98], [253, 231, 37], [230, 85, 13], [158, 1, 66],
                     [116, 196, 118], [107, 174, 214], [240, 240, 240], [82, 82, 82], [188, 128, 189]], dtype=np.uint8)


def render_segments(clusters, scale=8):
    """uint8 RGB image of the segment label map (the reference returns a matplotlib canvas of `plt.imshow(clusters)`,
    utils/attention_utils.py:266-276; same content, not the same pixels: matplotlib is not a dependency here)."""
    img = _PALETTE[np.asarray(clusters) % len(_PALETTE)]
    return np.repeat(np.repeat(img, scale, axis=0), scale, axis=1)


def render_token_maps(resized):
    """uint8 grey image with the region masks side by side (stands in for plot_attention_maps, :96-149, 334-335)."""
    tiles = [(m.clamp(0, 1) * 255).round().to(torch.uint8).cpu().numpy() for m in resized]
    return np.concatenate(tiles, axis=1)


def get_token_maps(selfattn_maps, crossattn_maps, n_maps, save_dir, width, height, obj_tokens, seed=0,
                   tokens_vis=None, preprocess=False, segment_threshold=0.3, num_segments=5, return_vis=False,
                   save_attn=False, device=None, return_clusters=False):
    """Same signature and return value as utils/attention_utils.py:233-341: a list of N masks
    [1, 4, height, width] fp32 (spans..., background) that sum to one per pixel."""
    from sklearn.cluster import SpectralClustering
    resolution = 32
    some = next(iter(selfattn_maps.values()))
    device = device or some.device
    aff = self_affinity(selfattn_maps, resolution).cpu().numpy()
    seed_everything(seed)
    sc = SpectralClustering(num_segments, affinity="precomputed", n_init=100, assign_labels="kmeans")
    clusters = sc.fit_predict(aff).reshape(resolution, resolution)
    cross = cross_maps_mean(crossattn_maps, resolution).cpu().numpy()

    normalized_span_maps = []
    for token_ids in obj_tokens:
        span = cross[:, :, np.asarray(token_ids.cpu())]
        norm = np.zeros_like(span)
        for i in range(span.shape[-1]):
            cur = span[:, :, i]
            norm[:, :, i] = (cur - np.abs(cur.min())) / (cur.max() - cur.min())   # sic, :302-303
        normalized_span_maps.append(norm)
    fg = [np.zeros([resolution, resolution]) for _ in normalized_span_maps]
    bg = np.zeros([resolution, resolution])
    for c in range(num_segments):
        cluster_mask = np.zeros_like(clusters)
        cluster_mask[clusters == c] = 1.0
        is_fg = False
        for norm, fg_map, token_ids in zip(normalized_span_maps, fg, obj_tokens):
            scores = [(cluster_mask * norm[:, :, i]).sum() / cluster_mask.sum() for i in range(len(token_ids))]
            if max(scores) > segment_threshold:
                fg_map += cluster_mask
                is_fg = True
        if not is_fg:
            bg += cluster_mask
    fg.append(bg)
    maps = torch.from_numpy(np.stack(fg)).to(device)                       # float64, as the reference
    resized = _resize(maps[:, None], (height, width))[:, 0].clamp(0, 1)     # (height, width) order, :325
    resized = resized / (resized.sum(0, True) + 1e-8)
    out = [m[None, None].repeat(1, 4, 1, 1).to(torch.float32) for m in resized]
    if return_clusters:   # (masks, int label image): extension used by the tests
        return out, clusters
    if return_vis:
        return out, render_segments(clusters), render_token_maps(resized)
    return out

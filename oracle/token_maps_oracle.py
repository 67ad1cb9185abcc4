"""TEST INFRASTRUCTURE — restatement of utils/attention_utils.py:233-341 get_token_maps (CPU, numpy/torch).

Spectral clustering is the third-party scikit-learn call of the reference (attention_utils.py:262-265,
pinned `scikit-learn==0.24.1` in environment.yaml; the installed version is used on both arms).
"""
import random

import numpy as np
import torch


def seed_everything(seed):
    """utils/richtext_utils.py:22-27."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def self_affinity(selfattn_maps, resolution=32):
    """attention_utils.py:241-256: keep the maps whose side == resolution, average over layers."""
    per_res = {8: [], 16: [], 32: [], 64: []}
    for attn_map in selfattn_maps.values():
        r = int(np.sqrt(attn_map.shape[1]).astype(int))
        if r != resolution:
            continue
        m = attn_map.reshape(1, r, r, r * r).permute([3, 0, 1, 2]).float()
        m = torch.nn.functional.interpolate(m, (resolution, resolution), mode="bicubic", antialias=True)
        per_res[r].append(m.permute([1, 2, 3, 0]).reshape(1, resolution ** 2, r ** 2))
    return torch.cat([torch.cat(v).mean(0).cpu() for v in per_res.values() if len(v) > 0], -1).numpy()


def cross_maps_mean(crossattn_maps, resolution=32):
    """attention_utils.py:281-291."""
    outs = []
    for attn_map in crossattn_maps.values():
        r = int(np.sqrt(attn_map.shape[1]).astype(int))
        m = attn_map.reshape(1, r, r, -1).permute([0, 3, 1, 2]).float()
        m = torch.nn.functional.interpolate(m, (resolution, resolution), mode="bicubic", antialias=True)
        outs.append(m.permute([0, 2, 3, 1]))
    return torch.cat(outs).mean(0).cpu().numpy()


def cluster(affinity, num_segments, seed, resolution=32):
    """attention_utils.py:261-265."""
    from sklearn.cluster import SpectralClustering
    seed_everything(seed)
    sc = SpectralClustering(num_segments, affinity="precomputed", n_init=100, assign_labels="kmeans")
    return sc.fit_predict(affinity).reshape(resolution, resolution)


def label_segments(clusters, cross_mean, obj_tokens, num_segments, segment_threshold):
    """attention_utils.py:296-323: per-span min/max normalisation (note `- abs(min)`), cluster scoring."""
    normalized_span_maps = []
    for token_ids in obj_tokens:
        span = cross_mean[:, :, token_ids.numpy()]
        norm = np.zeros_like(span)
        for i in range(span.shape[-1]):
            cur = span[:, :, i]
            norm[:, :, i] = (cur - np.abs(cur.min())) / (cur.max() - cur.min())
        normalized_span_maps.append(norm)
    fg = [np.zeros([clusters.shape[0], clusters.shape[1]]).squeeze() for _ in normalized_span_maps]
    bg = np.zeros([clusters.shape[0], clusters.shape[1]]).squeeze()
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
    return fg


def resize_and_normalise(token_maps, width, height):
    """attention_utils.py:325-337 (note the (height, width) order and the fp32 output dtype)."""
    r = torch.cat([torch.nn.functional.interpolate(torch.from_numpy(m).unsqueeze(0).unsqueeze(0), (height, width),
                                                   mode="bicubic", antialias=True)[0] for m in token_maps]).clamp(0, 1)
    r = r / (r.sum(0, True) + 1e-8)
    return [m.unsqueeze(0).unsqueeze(1).repeat([1, 4, 1, 1]).to(torch.float32) for m in r]


def get_token_maps(selfattn_maps, crossattn_maps, n_maps, save_dir, width, height, obj_tokens, seed=0,
                   tokens_vis=None, preprocess=False, segment_threshold=0.3, num_segments=5, return_vis=False,
                   save_attn=False, return_clusters=False):
    aff = self_affinity(selfattn_maps)
    clusters = cluster(aff, num_segments, seed)
    cross = cross_maps_mean(crossattn_maps)
    maps = label_segments(clusters, cross, obj_tokens, num_segments, segment_threshold)
    out = resize_and_normalise(maps, width, height)
    if return_clusters:
        return out, clusters
    return out

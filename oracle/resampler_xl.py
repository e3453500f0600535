"""Oracle for the de-tokenizer front: ResamplerXLV2 perceiver (fp32, CPU, functional).

Restates /root/reference/src/models/detokenizer/resampler.py: FeedForward 9-16, PerceiverAttention.forward 46-75,
AttentionPool2d.forward 89-116, ResamplerXLV2.forward 266-286.
PINNED: tests/golden/resampler_xl.pt holds outputs of the reference ResamplerXLV2 itself (make_golden.py).
"""
import math

import torch
import torch.nn.functional as F


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _heads(x, h):
    B, L, W = x.shape
    return x.reshape(B, L, h, W // h).transpose(1, 2)


def perceiver_attention(sd, p, x, lat, heads):
    xn, ln = _ln(sd, p + ".norm1", x), _ln(sd, p + ".norm2", lat)
    q = ln @ sd[p + ".to_q.weight"].t()
    k, v = (torch.cat([xn, ln], dim=1) @ sd[p + ".to_kv.weight"].t()).chunk(2, dim=-1)
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    w = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), dim=-1)   # (q d^-1/4)(k d^-1/4)^T == q k^T / sqrt(d)
    o = (w @ v).transpose(1, 2).reshape(lat.shape[0], lat.shape[1], -1)
    return o @ sd[p + ".to_out.weight"].t()


def attention_pool(sd, p, h, heads):
    """AttentionPool2d: query = mean token (+pos), keys/values = [mean | tokens] + pos; returns token 0 only."""
    t = torch.cat([h.mean(dim=1, keepdim=True), h], dim=1) + sd[p + ".positional_embedding"][None]
    q = t[:, :1] @ sd[p + ".q_proj.weight"].t() + sd[p + ".q_proj.bias"]
    k = t @ sd[p + ".k_proj.weight"].t() + sd[p + ".k_proj.bias"]
    v = t @ sd[p + ".v_proj.weight"].t() + sd[p + ".v_proj.bias"]
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    w = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), dim=-1)
    o = (w @ v).transpose(1, 2).reshape(h.shape[0], -1)
    return o @ sd[p + ".c_proj.weight"].t() + sd[p + ".c_proj.bias"]


def resampler_xl(sd, cfg, feats):
    """ResamplerXLV2.forward (normalize=False as in every shipped config): feats [B,n,embedding_dim] ->
    (prompt_embeds [B,num_queries,out1+out2], pooled [B,out2])."""
    sd = {k: v.float() for k, v in sd.items()}
    x = feats.float() @ sd["proj_in.weight"].t() + sd["proj_in.bias"]
    lat = sd["latents"].repeat(x.shape[0], 1, 1)
    for i in range(cfg["depth"]):
        lat = lat + perceiver_attention(sd, f"layers.{i}.0", x, lat, cfg["heads"])
        f = f"layers.{i}.1"
        lat = lat + F.gelu(_ln(sd, f + ".0", lat) @ sd[f + ".1.weight"].t()) @ sd[f + ".3.weight"].t()
    h = _ln(sd, "norm_out", lat)
    p1 = h @ sd["unet_proj_1.weight"].t() + sd["unet_proj_1.bias"]
    p2 = h @ sd["unet_proj_2.weight"].t() + sd["unet_proj_2.bias"]
    return torch.cat([p1, p2], dim=-1), attention_pool(sd, "unet_attnpool", h, cfg["heads"])

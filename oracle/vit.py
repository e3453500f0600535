"""Oracle for stage 1: Qwen ViT-bigG with attention pooling (fp32, CPU, functional over a state dict).

Restates /root/reference/src/models/tokenizer/qwen_visual.py:
  get_abs_pos 24-40, Resampler.forward 136-146, VisualAttention.forward 180-230, VisualAttentionBlock.forward 270-282,
  VisionTransformerWithAttnPool.forward 387-417.
PINNED: tests/golden/vit_small.pt and resampler_*.pt are outputs of the reference modules themselves
(tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file against them.
"""
import math

import torch
import torch.nn.functional as F


def resize_pos(table, n_tokens):
    """bicubic resize of a [g*g, C] position table to n_tokens = G*G rows (qwen_visual.py:24-40)."""
    src = int(math.isqrt(table.shape[0]))
    dst = int(math.isqrt(n_tokens))
    if src == dst:
        return table
    t = table.float().reshape(1, src, src, -1).permute(0, 3, 1, 2)
    t = F.interpolate(t, size=(dst, dst), mode="bicubic", align_corners=False)
    return t.permute(0, 2, 3, 1).reshape(dst * dst, -1)


def layer_norm(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def mha(q, k, v, heads):
    """softmax(q k^T / sqrt(d)) v with heads split contiguously; q [B,Nq,E], k/v [B,Nk,E]."""
    B, Nq, E = q.shape
    d = E // heads
    qh = q.reshape(B, Nq, heads, d).transpose(1, 2)
    kh = k.reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.reshape(B, -1, heads, d).transpose(1, 2)
    att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d), dim=-1)
    return (att @ vh).transpose(1, 2).reshape(B, Nq, E)


def resampler(sd, prefix, x, heads, eps):
    """qwen_visual.Resampler.forward (136-146): x [B,Nk,kv_dim] -> [B,Nq,E]."""
    E = sd[prefix + "query"].shape[1]
    if prefix + "kv_proj.weight" in sd:
        x = x @ sd[prefix + "kv_proj.weight"].t()
    kv = layer_norm(x, sd, prefix + "ln_kv", eps)
    pos = sd[prefix + "pos_embed"]
    qin = layer_norm(sd[prefix + "query"], sd, prefix + "ln_q", eps) + pos
    kin = kv + resize_pos(pos, kv.shape[1]).unsqueeze(0)
    wq, wk, wv = sd[prefix + "attn.in_proj_weight"].chunk(3, dim=0)
    bq, bk, bv = sd[prefix + "attn.in_proj_bias"].chunk(3, dim=0)
    q = (qin @ wq.t() + bq).unsqueeze(0).expand(x.shape[0], -1, -1)
    k = kin @ wk.t() + bk
    v = kv @ wv.t() + bv
    o = mha(q, k, v, heads)
    return o @ sd[prefix + "attn.out_proj.weight"].t() + sd[prefix + "attn.out_proj.bias"]


def vit_attention(sd, prefix, x, heads):
    """VisualAttention.forward (180-230): in_proj rows are head-major [h][q|k|v]; x [B,S,E]."""
    B, S, E = x.shape
    d = E // heads
    qkv = x @ sd[prefix + "in_proj.weight"].t() + sd[prefix + "in_proj.bias"]
    qkv = qkv.reshape(B, S, heads, 3, d)
    q, k, v = (qkv[:, :, :, i].transpose(1, 2) for i in range(3))  # [B,H,S,d]
    att = torch.softmax((q / math.sqrt(d)) @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, E)
    return o @ sd[prefix + "out_proj.weight"].t() + sd[prefix + "out_proj.bias"]


def vit_forward(sd, x, heads, eps=1e-6, return_tokens=False):
    """VisionTransformerWithAttnPool.forward (387-417): x [N,3,H,W] -> [N,n_queries,output_dim] (fp32)."""
    sd = {k: v.float() for k, v in sd.items()}
    x = x.float()
    patch = sd["conv1.weight"].shape[-1]
    tok = F.conv2d(x, sd["conv1.weight"], stride=patch).flatten(2).transpose(1, 2)  # [N,S,W]
    tok = tok + resize_pos(sd["positional_embedding"], tok.shape[1])
    h = layer_norm(tok, sd, "ln_pre", eps)
    layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        h = h + vit_attention(sd, p + "attn.", layer_norm(h, sd, p + "ln_1", eps), heads)
        m = F.gelu(layer_norm(h, sd, p + "ln_2", eps) @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"])
        h = h + m @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"]
    if return_tokens:
        return h
    out_dim = sd["proj"].shape[0]
    z = resampler(sd, "attn_pool.", h, out_dim // 128, eps)
    z = layer_norm(z, sd, "ln_post", eps)
    return z @ sd["proj"]


def vit_down(feats, k=4):
    """token mean-pool 256 -> 64 ('vit_down'; adapter_modules.py:112-115, seed_x.py:103-106)."""
    return F.avg_pool1d(feats.transpose(1, 2), kernel_size=k, stride=k).transpose(1, 2)

"""Oracle for stage 2: LLaMA causal LM forward with KV cache, the greedy loop with the image-token logits processor, and
ContinuousLVLM.generate (fp32, CPU, functional over HF-named state dicts).

Restates
  /root/reference/src/models/mllm/modeling_llama_xformer.py  (RoPE 97-149, LlamaAttention.forward 193-244, LlamaMLP 152-167,
      LlamaDecoderLayer.forward 261-313, LlamaModel.forward 477-609, LlamaForCausalLM.forward 643-746),
  /root/reference/src/models/mllm/generation.py:19-31 (AutoImageTokenGenerationProcessor.__call__),
  /root/reference/src/models/mllm/seed_x.py:130-223 (ContinuousLVLM.generate),
  /root/reference/src/models/mllm/peft_models.py:62-82 + PEFT 0.4.0 (vendored: /root/reference/proj/peft) LoRA forward, lora.py:808-832,
  transformers==4.30.2 GenerationMixin.greedy_search as used at seed_x.py:184-189 (THIRD-PARTY, absent:
  the installed transformers 5.5.0 cannot drive the reference's xformers model class, but tests/test_hf_generate_pin_cpu.py pins this loop and
  the forward below against HF's own LlamaForCausalLM.generate called the way seed_x.py:184-189 calls it: ids, EOS stop, harvested hidden rows).
PINNED: forward logits / hidden states / KV and the logits processor against golden vectors produced by the reference
modules themselves (tests/golden/llama_tiny.pt, make_golden.py); the greedy loop against the same loop run around the
reference forward + the reference's own processor class; the LoRA / vocabulary-growth path against the reference's
get_peft_model_with_resize_embedding over its vendored PEFT (tests/golden/llama_lora_tiny.pt).
"""
import math

import torch
import torch.nn.functional as F

from . import vit as ovit


def rms_norm(x, w, eps):
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def rope(x, pos, base=10000.0):
    """x [T,H,d], pos [T]: x*cos + rotate_half(x)*sin with frequencies duplicated over both halves (97-149)."""
    d = x.shape[-1]
    inv = 1.0 / (base ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    ang = pos.float()[:, None] * inv[None, :]
    emb = torch.cat([ang, ang], dim=-1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]
    rot = torch.cat([-x[..., d // 2:], x[..., : d // 2]], dim=-1)
    return x * cos + rot * sin


def resize_embeddings(sd, new_vocab):
    """get_peft_model_with_resize_embedding's vocabulary growth (/root/reference/src/models/mllm/peft_models.py:62-82): appended input
    rows = mean of the old input rows; appended output rows = 3 x the mean of the old output rows."""
    sd = dict(sd)
    for key, gain in (("model.embed_tokens.weight", 1.0), ("lm_head.weight", 3.0)):
        w = sd[key].float()
        extra = new_vocab - w.shape[0]
        if extra > 0:
            sd[key] = torch.cat([w, (w.mean(dim=0, keepdim=True) * gain).expand(extra, -1)], dim=0)
    return sd


def lora_linear(sd, lora, name, x):
    """PEFT 0.4.0 ``Linear.forward`` un-merged, eval mode (/root/reference/proj/peft/src/peft/tuners/lora.py:808-832):
    ``F.linear(x, W) + lora_B(lora_A(x)) * (lora_alpha / r)``.  lora = dict(scaling=, sd={'<name>.lora_A.weight', '<name>.lora_B.weight'})."""
    y = x @ sd[name + ".weight"].float().t()
    if lora is not None and name + ".lora_A.weight" in lora["sd"]:
        a, b = lora["sd"][name + ".lora_A.weight"].float(), lora["sd"][name + ".lora_B.weight"].float()
        y = y + ((x @ a.t()) @ b.t()) * lora["scaling"]
    return y


def llama_forward(sd, cfg, x, pos0, cache, lora=None):
    """x: [T, D] input embeddings of positions pos0..pos0+T-1; cache: list per layer of (K [t,H,d], V [t,H,d]) or None.
    Returns (logits [T,V], last_hidden post-final-norm [T,D], new cache).  Causal within the new tokens, full over the past.
    lora: optional un-merged adapters (see lora_linear)."""
    lin = lambda name, t: lora_linear(sd, lora, name, t)  # noqa: E731
    D, H = cfg["hidden"], cfg["heads"]
    d = D // H
    T = x.shape[0]
    pos = torch.arange(pos0, pos0 + T)
    new_cache = []
    h = x.float()
    for i in range(cfg["layers"]):
        p = f"model.layers.{i}."
        n = rms_norm(h, sd[p + "input_layernorm.weight"], cfg["eps"])
        q = lin(p + "self_attn.q_proj", n).reshape(T, H, d)
        k = lin(p + "self_attn.k_proj", n).reshape(T, H, d)
        v = lin(p + "self_attn.v_proj", n).reshape(T, H, d)
        q, k = rope(q, pos), rope(k, pos)
        if cache is not None and cache[i] is not None:
            k = torch.cat([cache[i][0], k], dim=0)
            v = torch.cat([cache[i][1], v], dim=0)
        new_cache.append((k, v))
        S = k.shape[0]
        att = torch.einsum("thd,shd->hts", q, k) / math.sqrt(d)
        mask = torch.arange(S)[None, :] > (pos0 + torch.arange(T))[:, None]   # key s visible to query t iff s <= pos0 + t
        att = att.masked_fill(mask[None], float("-inf")).softmax(-1)
        o = torch.einsum("hts,shd->thd", att, v).reshape(T, D)
        h = h + lin(p + "self_attn.o_proj", o)
        n = rms_norm(h, sd[p + "post_attention_layernorm.weight"], cfg["eps"])
        g = F.silu(lin(p + "mlp.gate_proj", n)) * lin(p + "mlp.up_proj", n)
        h = h + lin(p + "mlp.down_proj", g)
    hn = rms_norm(h, sd["model.norm.weight"], cfg["eps"])
    return hn @ sd["lm_head.weight"].t(), hn, new_cache


def image_token_processor(img_ids, last_id, scores):
    """AutoImageTokenGenerationProcessor.__call__ (generation.py:19-31) for one row: img_ids = ids of <img><img_0>..<img_n-1></img>."""
    scores = scores.clone()
    if last_id in img_ids[:-1]:
        nxt = img_ids[img_ids.index(last_id) + 1]
        scores[nxt] = scores.max() + 10.0
    else:
        scores[torch.tensor(img_ids[1:], dtype=torch.long)] = 0.0
    return scores


def greedy_generate(sd, cfg, input_ids, inputs_embeds, img_ids, max_new_tokens, eos_id=None):
    """transformers-4.30 greedy_search as called at seed_x.py:184-189 (B.1): step 0 consumes inputs_embeds [P,D]; later steps the
    last id.  Returns (generated ids list, last_hidden rows [n_generated-1, D] = post-norm state of the position that consumed
    generated token j, as sliced at seed_x.py:196-197)."""
    emb = sd["model.embed_tokens.weight"]
    P = inputs_embeds.shape[0]
    logits, hid, cache = llama_forward(sd, cfg, inputs_embeds, 0, None)
    seq = list(input_ids)
    gen, hiddens = [], []
    for step in range(max_new_tokens):
        s = image_token_processor(img_ids, seq[-1], logits[-1])
        nxt = int(torch.argmax(s))
        gen.append(nxt)
        seq.append(nxt)
        if (eos_id is not None and nxt == eos_id) or step == max_new_tokens - 1:
            break
        logits, hid, cache = llama_forward(sd, cfg, emb[nxt][None].float(), P + step, cache)
        hiddens.append(hid[-1])
    return gen, (torch.stack(hiddens) if hiddens else torch.zeros(0, cfg["hidden"]))


def lvlm_generate(llm_sd, agent_sd, cfg, tok, input_ids, image_embeds, ids_cmp_mask, embeds_cmp_mask, patch_positions, max_new_tokens,
                  num_img_tokens=64, eos_id=None):
    """ContinuousLVLM.generate (seed_x.py:130-223), batch 1.  image_embeds [N,256,vit_dim] or None."""
    D = cfg["hidden"]
    ids = list(input_ids)
    x = llm_sd["model.embed_tokens.weight"][torch.tensor(ids)].float().clone()
    if image_embeds is not None:
        lm = ovit.resampler(agent_sd, "input_resampler.", image_embeds.float(), 32 if D == 5120 else cfg.get("resampler_heads", 2), 1e-5)
        pp = torch.cat([patch_positions, 1 - patch_positions], dim=-1).float() / 2
        lm = lm + (pp @ agent_sd["patch_pos_embed"].float()).unsqueeze(1)
        x[ids_cmp_mask] = lm[embeds_cmp_mask].reshape(-1, D)
    img_ids = tok.encode("".join(["<img>"] + ["<img_{:05d}>".format(i) for i in range(num_img_tokens)] + ["</img>"]), add_special_tokens=False)
    gen, hid = greedy_generate(llm_sd, cfg, ids, x, img_ids, max_new_tokens, eos_id)
    eoi, boi = tok.encode("</img>")[0], tok.encode("<img>")[0]
    g = torch.tensor(gen)
    eoi_idx = torch.where(g == eoi)[0].tolist()
    feats = None
    text_mask = torch.ones_like(g, dtype=torch.bool)
    if eoi_idx:
        rows = [hid[e - num_img_tokens:e] for e in eoi_idx]
        for e in eoi_idx:
            text_mask[e - num_img_tokens:e] = False
        out_heads = agent_sd["output_resampler.query"].shape[1] // 128 if D == 5120 else cfg.get("resampler_heads", 2)
        feats = ovit.resampler(agent_sd, "output_resampler.", torch.stack(rows), out_heads, 1e-5)
    text_mask[g == boi] = False
    return dict(ids=gen, text=tok.decode(g[text_mask]), has_img_output=bool(eoi_idx), img_gen_feat=feats, num_gen_imgs=len(eoi_idx))

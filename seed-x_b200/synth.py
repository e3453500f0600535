"""Deterministic synthetic weights / inputs (no checkpoints or tokenizer files exist offline).

Every tensor is a pure function of (seed, name, shape): values come from a torch CPU generator seeded by a hash of the
name, so the CPU oracle, the golden-vector script and the CUDA path all see identical parameters.  Tensor names and
shapes follow the reference modules' state dicts.  Linear/conv weights use std = fan_in^-0.5
so branch outputs are O(1) and softmaxes are non-uniform.
"""
import hashlib
import math
from collections import OrderedDict

import torch

SEED = 1234


def _gen(name, seed):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & ((1 << 62) - 1))
    return g


_DEVICE = "cpu"


def set_device(device):
    """'cpu' (default): bit-reproducible fp32 tensors shared by oracle, goldens and CUDA path.  'cuda': tensors are drawn directly
    on the GPU in fp16 (a different stream of values) — used to materialise the 13 B-parameter benchmark models in seconds."""
    global _DEVICE
    _DEVICE = device


def randn(name, shape, std=1.0, mean=0.0, seed=SEED):
    if _DEVICE != "cpu":
        h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
        g = torch.Generator(device=_DEVICE)
        g.manual_seed(int.from_bytes(h[:8], "little") & ((1 << 62) - 1))
        t = torch.randn(tuple(shape), generator=g, device=_DEVICE, dtype=torch.float16)
        return t.mul_(std).add_(mean) if (std != 1.0 or mean != 0.0) else t
    t = torch.randn(tuple(shape), generator=_gen(name, seed), dtype=torch.float32)
    # fp16-representable values: the reference checkpoints are fp16 tensors, so the oracle (fp32 math) and the CUDA path
    # (fp16 operands) start from bit-identical parameters and inputs
    return (t * std + mean).half().float()


def sincos_2d(embed_dim, grid):
    """Fixed 2-D sin/cos table of the Resampler (restates qwen_visual.py:44-91): first half encodes the x (width)
    coordinate, second half y; each half = [sin(p*w_i) | cos(p*w_i)], w_i = 10000^(-i/(D/4))."""
    q = embed_dim // 4
    omega = 1.0 / (10000.0 ** (torch.arange(q, dtype=torch.float32) / float(q)))
    ys, xs = torch.meshgrid(torch.arange(grid, dtype=torch.float32), torch.arange(grid, dtype=torch.float32), indexing="ij")

    def enc(p):
        o = p.reshape(-1, 1) * omega.reshape(1, -1)
        return torch.cat([torch.sin(o), torch.cos(o)], dim=1)

    return torch.cat([enc(xs), enc(ys)], dim=1)  # kept fp32: a buffer the reference rebuilds at construction


def _norm(sd, prefix, dim, bias=True):
    sd[prefix + ".weight"] = randn(prefix + ".weight", (dim,), 0.1, 1.0)
    if bias:
        sd[prefix + ".bias"] = randn(prefix + ".bias", (dim,), 0.1)


def _linear(sd, prefix, out_f, in_f, bias=True, wname="weight"):
    sd[f"{prefix}.{wname}"] = randn(f"{prefix}.{wname}", (out_f, in_f), in_f ** -0.5)
    if bias:
        sd[prefix + ".bias"] = randn(prefix + ".bias", (out_f,), 0.1)


def resampler_state_dict(prefix, grid, embed_dim, kv_dim, sd=None):
    """qwen_visual.Resampler parameters (attn_pool / input_resampler / output_resampler)."""
    sd = OrderedDict() if sd is None else sd
    p = prefix
    sd[p + "pos_embed"] = sincos_2d(embed_dim, grid)
    sd[p + "query"] = randn(p + "query", (grid * grid, embed_dim), 1.0)
    if kv_dim != embed_dim:
        sd[p + "kv_proj.weight"] = randn(p + "kv_proj.weight", (embed_dim, kv_dim), kv_dim ** -0.5)
    sd[p + "attn.in_proj_weight"] = randn(p + "attn.in_proj_weight", (3 * embed_dim, embed_dim), embed_dim ** -0.5)
    sd[p + "attn.in_proj_bias"] = randn(p + "attn.in_proj_bias", (3 * embed_dim,), 0.1)
    _linear(sd, p + "attn.out_proj", embed_dim, embed_dim)
    _norm(sd, p + "ln_q", embed_dim)
    _norm(sd, p + "ln_kv", embed_dim)
    return sd


def vit_state_dict(width=1664, layers=48, heads=16, mlp_width=8192, output_dim=4096, n_queries=256, patch=14):
    """VisionTransformerWithAttnPool parameters (qwen_visual.py:325-385)."""
    sd = OrderedDict()
    sd["positional_embedding"] = randn("positional_embedding", (256, width), width ** -0.5)
    sd["proj"] = randn("proj", (output_dim, output_dim), output_dim ** -0.5)
    sd["conv1.weight"] = randn("conv1.weight", (width, 3, patch, patch), (3 * patch * patch) ** -0.5)
    _norm(sd, "ln_pre", width)
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        _norm(sd, p + "ln_1", width)
        _norm(sd, p + "ln_2", width)
        _linear(sd, p + "attn.in_proj", 3 * width, width)
        _linear(sd, p + "attn.out_proj", width, width)
        _linear(sd, p + "mlp.c_fc", mlp_width, width)
        _linear(sd, p + "mlp.c_proj", width, mlp_width)
    resampler_state_dict("attn_pool.", int(math.isqrt(n_queries)), output_dim, width, sd)
    _norm(sd, "ln_post", output_dim)
    return sd


def image(name, n, size, seed=SEED):
    """CLIP-normalised-looking image batch [n,3,size,size] fp32."""
    return randn(name, (n, 3, size, size), 1.0, seed=seed)


# ----------------------------------------------------------------------------------------------------------------------
# SDXL UNet / VAE parameters (diffusers state-dict names)
# ----------------------------------------------------------------------------------------------------------------------
SDXL_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=2,
                 down_attn=(False, True, True), transformer_layers=(1, 2, 10), heads=(5, 10, 20), cross_attention_dim=2048,
                 time_embed_dim=1280, addition_time_embed_dim=256, text_embed_dim=1280, groups=32)
SDXL_VAE = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, groups=32, scaling_factor=0.13025)
# scaled-down configs with the same topology for oracle-sized parity tests
TINY_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 256), layers_per_block=2,
                 down_attn=(False, True, True), transformer_layers=(1, 1, 2), heads=(1, 2, 4), cross_attention_dim=128,
                 time_embed_dim=256, addition_time_embed_dim=32, text_embed_dim=64, groups=32)
TINY_VAE = dict(block_out_channels=(32, 64, 64, 64), layers_per_block=1, latent_channels=4, groups=32, scaling_factor=0.13025)


def _conv(sd, p, cout, cin, k, gain=1.0):
    sd[p + ".weight"] = randn(p + ".weight", (cout, cin, k, k), gain * (cin * k * k) ** -0.5)
    sd[p + ".bias"] = randn(p + ".bias", (cout,), 0.05)


def _lin(sd, p, out_f, in_f, bias=True, gain=1.0):
    sd[p + ".weight"] = randn(p + ".weight", (out_f, in_f), gain * in_f ** -0.5)
    if bias:
        sd[p + ".bias"] = randn(p + ".bias", (out_f,), 0.05)


def _resnet(sd, p, cin, cout, temb):
    _norm(sd, p + ".norm1", cin)
    _conv(sd, p + ".conv1", cout, cin, 3)
    if temb:
        _lin(sd, p + ".time_emb_proj", cout, temb)
    _norm(sd, p + ".norm2", cout)
    _conv(sd, p + ".conv2", cout, cout, 3, gain=0.5)
    if cin != cout:
        _conv(sd, p + ".conv_shortcut", cout, cin, 1)


def _transformer(sd, p, c, depth, ctx_dim):
    _norm(sd, p + ".norm", c)
    _lin(sd, p + ".proj_in", c, c)
    for k in range(depth):
        b = f"{p}.transformer_blocks.{k}"
        for n in ("norm1", "norm2", "norm3"):
            _norm(sd, f"{b}.{n}", c)
        for a, kd in (("attn1", c), ("attn2", ctx_dim)):
            _lin(sd, f"{b}.{a}.to_q", c, c, bias=False)
            _lin(sd, f"{b}.{a}.to_k", c, kd, bias=False)
            _lin(sd, f"{b}.{a}.to_v", c, kd, bias=False)
            _lin(sd, f"{b}.{a}.to_out.0", c, c, gain=0.5)
        _lin(sd, f"{b}.ff.net.0.proj", 8 * c, c)
        _lin(sd, f"{b}.ff.net.2", c, 4 * c, gain=0.5)
    _lin(sd, p + ".proj_out", c, c, gain=0.5)


def unet_state_dict(cfg, prefix=""):
    sd = OrderedDict()
    boc = cfg["block_out_channels"]
    te = cfg["time_embed_dim"]
    nb = len(boc)
    _conv(sd, "conv_in", boc[0], cfg["in_channels"], 3)
    _lin(sd, "time_embedding.linear_1", te, boc[0])
    _lin(sd, "time_embedding.linear_2", te, te)
    _lin(sd, "add_embedding.linear_1", te, cfg["text_embed_dim"] + 6 * cfg["addition_time_embed_dim"])
    _lin(sd, "add_embedding.linear_2", te, te)
    ch = boc[0]
    for i in range(nb):
        for j in range(cfg["layers_per_block"]):
            _resnet(sd, f"down_blocks.{i}.resnets.{j}", ch, boc[i], te)
            ch = boc[i]
            if cfg["down_attn"][i]:
                _transformer(sd, f"down_blocks.{i}.attentions.{j}", ch, cfg["transformer_layers"][i], cfg["cross_attention_dim"])
        if i < nb - 1:
            _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", ch, ch, 3)
    _resnet(sd, "mid_block.resnets.0", ch, ch, te)
    _transformer(sd, "mid_block.attentions.0", ch, cfg["transformer_layers"][-1], cfg["cross_attention_dim"])
    _resnet(sd, "mid_block.resnets.1", ch, ch, te)
    rev = list(reversed(boc))
    for i in range(nb):
        r = nb - 1 - i
        out_c = rev[i]
        prev = rev[i - 1] if i > 0 else rev[0]
        inp = rev[min(i + 1, nb - 1)]
        for j in range(cfg["layers_per_block"] + 1):
            skip = inp if j == cfg["layers_per_block"] else out_c
            rin = prev if j == 0 else out_c
            _resnet(sd, f"up_blocks.{i}.resnets.{j}", rin + skip, out_c, te)
            if cfg["down_attn"][r]:
                _transformer(sd, f"up_blocks.{i}.attentions.{j}", out_c, cfg["transformer_layers"][r], cfg["cross_attention_dim"])
        if i < nb - 1:
            _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    _norm(sd, "conv_norm_out", boc[0])
    _conv(sd, "conv_out", cfg["out_channels"], boc[0], 3)
    if prefix:
        sd = OrderedDict((prefix + k, v) for k, v in sd.items())
    return sd


def _vae_attn(sd, p, c):
    _norm(sd, p + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v"):
        _lin(sd, f"{p}.{n}", c, c)
    _lin(sd, p + ".to_out.0", c, c, gain=0.5)


def vae_state_dict(cfg):
    sd = OrderedDict()
    boc = cfg["block_out_channels"]
    L = cfg["latent_channels"]
    # encoder
    _conv(sd, "encoder.conv_in", boc[0], 3, 3)
    ch = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg["layers_per_block"]):
            _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", ch, c, 0)
            ch = c
        if i < len(boc) - 1:
            _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", ch, ch, 3)
    _resnet(sd, "encoder.mid_block.resnets.0", ch, ch, 0)
    _vae_attn(sd, "encoder.mid_block.attentions.0", ch)
    _resnet(sd, "encoder.mid_block.resnets.1", ch, ch, 0)
    _norm(sd, "encoder.conv_norm_out", ch)
    _conv(sd, "encoder.conv_out", 2 * L, ch, 3)
    _conv(sd, "quant_conv", 2 * L, 2 * L, 1)
    # decoder
    _conv(sd, "post_quant_conv", L, L, 1)
    rev = list(reversed(boc))
    _conv(sd, "decoder.conv_in", rev[0], L, 3)
    ch = rev[0]
    _resnet(sd, "decoder.mid_block.resnets.0", ch, ch, 0)
    _vae_attn(sd, "decoder.mid_block.attentions.0", ch)
    _resnet(sd, "decoder.mid_block.resnets.1", ch, ch, 0)
    for i, c in enumerate(rev):
        for j in range(cfg["layers_per_block"] + 1):
            _resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", ch, c, 0)
            ch = c
        if i < len(rev) - 1:
            _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    _norm(sd, "decoder.conv_norm_out", ch)
    _conv(sd, "decoder.conv_out", 3, ch, 3)
    return sd


# ----------------------------------------------------------------------------------------------------------------------
# LLaMA / agent parameters and a synthetic tokenizer (no tokenizer files exist offline)
# ----------------------------------------------------------------------------------------------------------------------
LLAMA_13B = dict(vocab=32330, hidden=5120, layers=40, heads=40, ffn=13824, eps=1e-5)
TINY_LLAMA = dict(vocab=1024, hidden=256, layers=2, heads=2, ffn=512, eps=1e-5)


def llama_state_dict(cfg):
    """LlamaForCausalLM parameters (modeling_llama_xformer.py), HF key names."""
    sd = OrderedDict()
    D, FF, V = cfg["hidden"], cfg["ffn"], cfg["vocab"]
    d = D // cfg["heads"]
    sd["model.embed_tokens.weight"] = randn("model.embed_tokens.weight", (V, D), 1.0)
    for i in range(cfg["layers"]):
        p = f"model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[p + f"self_attn.{n}.weight"] = randn(p + f"self_attn.{n}.weight", (D, D), D ** -0.5)
        sd[p + "self_attn.rotary_emb.inv_freq"] = 1.0 / (10000.0 ** (torch.arange(0, d, 2).float() / d))
        sd[p + "mlp.gate_proj.weight"] = randn(p + "mlp.gate_proj.weight", (FF, D), D ** -0.5)
        sd[p + "mlp.up_proj.weight"] = randn(p + "mlp.up_proj.weight", (FF, D), D ** -0.5)
        sd[p + "mlp.down_proj.weight"] = randn(p + "mlp.down_proj.weight", (D, FF), FF ** -0.5)
        sd[p + "input_layernorm.weight"] = randn(p + "input_layernorm.weight", (D,), 0.1, 1.0)
        sd[p + "post_attention_layernorm.weight"] = randn(p + "post_attention_layernorm.weight", (D,), 0.1, 1.0)
    sd["model.norm.weight"] = randn("model.norm.weight", (D,), 0.1, 1.0)
    sd["lm_head.weight"] = randn("lm_head.weight", (V, D), D ** -0.5)
    return sd


def lora_fixture(shapes):
    """fine-tuned tensors of the LoRA fixture (tests/golden/llama_lora_tiny.pt): a pure function of the PEFT key and its shape —
    lora_A ~ in^-0.5, lora_B ~ 0.15 r^-0.5 (with alpha/r = 2 the dense update is ~0.3 of the base weight's scale: clearly visible in the
    logits, yet a fine-tune, not a new network), modules_to_save norm weights ~ N(1, 0.1) like the base norms."""
    out = {}
    for k, shp in shapes.items():
        if ".lora_A." in k:
            out[k] = randn("lora:" + k, shp, std=shp[1] ** -0.5)
        elif ".lora_B." in k:
            out[k] = randn("lora:" + k, shp, std=0.15 * shp[1] ** -0.5)
        else:
            out[k] = randn("lora:" + k, shp, std=0.1, mean=1.0)
    return out


def agent_state_dict(llm_hidden, vit_dim, sd=None):
    """ContinuousLVLM parameters besides the LLM (seed_x.py:22-45): input/output Resampler + patch_pos_embed."""
    sd = OrderedDict() if sd is None else sd
    resampler_state_dict("input_resampler.", 8, llm_hidden, vit_dim, sd)
    resampler_state_dict("output_resampler.", 8, vit_dim, llm_hidden, sd)
    sd["patch_pos_embed"] = randn("patch_pos_embed", (4, llm_hidden), llm_hidden ** -0.5)
    return sd


class SynthTokenizer:
    """Stand-in for the LlamaTokenizer with the reference's added special tokens: ids [0, base) are text,
    then <img>, </img>, <patch>, </patch>, <box_start>, <box_end>, <img_00000..00099>, <loc-0..223>."""
    SPECIALS = ["<img>", "</img>", "<patch>", "</patch>", "<box_start>", "<box_end>"]

    def __init__(self, vocab=32330, n_img=100, n_loc=224):
        n_special = len(self.SPECIALS) + n_img + n_loc
        self.base = vocab - n_special
        assert self.base > 3
        self.vocab = vocab
        self.bos_token_id, self.eos_token_id, self.pad_token_id = 1, 2, 0
        self.tok2id = {s: self.base + i for i, s in enumerate(self.SPECIALS)}
        for i in range(n_img):
            self.tok2id["<img_{:05d}>".format(i)] = self.base + len(self.SPECIALS) + i
        for i in range(n_loc):
            self.tok2id["<loc-{}>".format(i)] = self.base + len(self.SPECIALS) + n_img + i
        self.id2tok = {v: k for k, v in self.tok2id.items()}

    def encode(self, text, add_special_tokens=False):
        """special-token strings map to their ids; any other character maps to a deterministic text id in [3, base)."""
        ids, i = [], 0
        if add_special_tokens:
            ids.append(self.bos_token_id)
        while i < len(text):
            if text[i] == "<":
                j = text.find(">", i)
                if j > 0 and text[i:j + 1] in self.tok2id:
                    ids.append(self.tok2id[text[i:j + 1]])
                    i = j + 1
                    continue
            ids.append(3 + (ord(text[i]) * 2654435761 % (self.base - 3)))
            i += 1
        return ids

    def __call__(self, text, return_tensors=None):
        ids = self.encode(text, add_special_tokens=True)

        class _R:
            pass

        r = _R()
        r.input_ids = torch.tensor([ids]) if return_tensors == "pt" else ids
        return r

    def decode(self, ids, skip_special_tokens=False):
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        return "".join(self.id2tok.get(int(i), f"[{int(i)}]") for i in ids)


# ----------------------------------------------------------------------------------------------------------------------
# ResamplerXLV2 (de-tokenizer front, resampler.py:226-286)
# ----------------------------------------------------------------------------------------------------------------------
RESAMPLER_XL = dict(dim=1024, depth=4, dim_head=64, heads=16, num_queries=64, embedding_dim=4096, output1_dim=768, output2_dim=1280, ff_mult=4)
TINY_RESAMPLER_XL = dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=64, embedding_dim=320, output1_dim=96, output2_dim=160, ff_mult=4)


def resampler_xl_state_dict(cfg, prefix=""):
    sd = OrderedDict()
    D, inner = cfg["dim"], cfg["dim_head"] * cfg["heads"]
    sd["latents"] = randn(prefix + "latents", (1, cfg["num_queries"], D), D ** -0.5 * 4)
    _lin(sd, "proj_in", D, cfg["embedding_dim"])
    _norm(sd, "norm_out", D)
    for i in range(cfg["depth"]):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        _norm(sd, a + ".norm1", D)
        _norm(sd, a + ".norm2", D)
        _lin(sd, a + ".to_q", inner, D, bias=False)
        _lin(sd, a + ".to_kv", 2 * inner, D, bias=False)
        _lin(sd, a + ".to_out", D, inner, bias=False, gain=0.5)
        _norm(sd, f + ".0", D)
        _lin(sd, f + ".1", int(D * cfg["ff_mult"]), D, bias=False)
        _lin(sd, f + ".3", D, int(D * cfg["ff_mult"]), bias=False, gain=0.5)
    _lin(sd, "unet_proj_1", cfg["output1_dim"], D)
    _lin(sd, "unet_proj_2", cfg["output2_dim"], D)
    sd["unet_attnpool.positional_embedding"] = randn("unet_attnpool.positional_embedding", (cfg["num_queries"] + 1, D), D ** -0.5)
    for n in ("k_proj", "q_proj", "v_proj"):
        _lin(sd, "unet_attnpool." + n, D, D)
    _lin(sd, "unet_attnpool.c_proj", cfg["output2_dim"], D)
    if prefix:
        sd = OrderedDict((prefix + k, v) for k, v in sd.items())
    return sd

"""Oracle for stage 3: SDXL UNet2DConditionModel, AutoencoderKL (VAE) and EulerDiscreteScheduler (fp32, CPU).

PARITY UNPINNED: the arithmetic lives in the third-party dependency diffusers==0.25.0 (requirements.txt:4), which is
neither vendored under /root/reference nor installed here, and the reference holds no test or golden vector for it.
This file restates the published diffusers-0.25.0 algorithms from the specification in SURVEY.md Appendix B.2
(UNet block walk, ResnetBlock2D, Transformer2DModel/BasicTransformerBlock, VAE encoder/decoder, Euler scheduler) and is
anchored on the reference's own call sites:
  src/models/detokenizer/pipeline_stable_diffusion_xl_t2i_edit.py:474-566, 823-994 (edit loop)
  src/models/detokenizer/adapter_modules.py:132-169 (t2i call into StableDiffusionXLPipeline)
State-dict keys are the diffusers names, so real checkpoints load unchanged.

PINNED around that gap: `edit_sample` (3-way CFG order, un-scaled image latents, sigma-space combine, Euler walk, time ids) reproduces at 2e-5 the
latents and image of the reference's OWN SDXLAdapterWithLatentImage.generate + StableDiffusionXLText2ImageAndEditPipeline.__call__ when those are
handed UNet / VAE / scheduler objects backed by this file (tests/golden/edit_adapter_tiny.pt, make_golden.py::golden_edit_adapter).  What stays
unpinned is the arithmetic INSIDE those three diffusers objects and the stock 2-way t2i pipeline.
"""
import math

import torch
import torch.nn.functional as F

SDXL_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=2,
                 down_attn=(False, True, True), transformer_layers=(1, 2, 10), heads=(5, 10, 20), cross_attention_dim=2048,
                 time_embed_dim=1280, addition_time_embed_dim=256, text_embed_dim=1280, groups=32)
SDXL_VAE = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, groups=32, scaling_factor=0.13025)


# ------------------------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t.float().reshape(-1, 1) * freqs.reshape(1, -1)
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=1)


def lin(sd, p, x):
    b = sd.get(p + ".bias")
    return F.linear(x, sd[p + ".weight"], b)


def conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def gn(sd, p, x, groups, eps):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def resnet(sd, p, x, emb, groups, eps):
    """ResnetBlock2D: GN-SiLU-conv3x3 (+ time proj) GN-SiLU-conv3x3, 1x1 shortcut when channels change."""
    h = conv(sd, p + ".conv1", F.silu(gn(sd, p + ".norm1", x, groups, eps)))
    if emb is not None:
        h = h + lin(sd, p + ".time_emb_proj", F.silu(emb))[:, :, None, None]
    h = conv(sd, p + ".conv2", F.silu(gn(sd, p + ".norm2", h, groups, eps)))
    if (p + ".conv_shortcut.weight") in sd:
        x = conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention(sd, p, x, ctx, heads):
    """diffusers Attention (AttnProcessor2_0): q from x, k/v from ctx; scale 1/sqrt(d)."""
    B, S, C = x.shape
    q, k, v = lin(sd, p + ".to_q", x), lin(sd, p + ".to_k", ctx), lin(sd, p + ".to_v", ctx)
    d = C // heads
    q = q.reshape(B, S, heads, d).transpose(1, 2)
    k = k.reshape(B, -1, heads, d).transpose(1, 2)
    v = v.reshape(B, -1, heads, d).transpose(1, 2)
    o = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1) @ v
    return lin(sd, p + ".to_out.0", o.transpose(1, 2).reshape(B, S, C))


def transformer_block(sd, p, x, ctx, heads):
    """BasicTransformerBlock: self-attn, cross-attn, GEGLU feed-forward, each pre-LN with residual."""
    n1 = ln(sd, p + ".norm1", x)
    x = x + attention(sd, p + ".attn1", n1, n1, heads)
    x = x + attention(sd, p + ".attn2", ln(sd, p + ".norm2", x), ctx, heads)
    hg = lin(sd, p + ".ff.net.0.proj", ln(sd, p + ".norm3", x))
    h, g = hg.chunk(2, dim=-1)
    return x + lin(sd, p + ".ff.net.2", h * F.gelu(g))


def transformer2d(sd, p, x, ctx, heads, depth, groups):
    """Transformer2DModel with use_linear_projection=True."""
    B, C, H, W = x.shape
    h = gn(sd, p + ".norm", x, groups, 1e-6).permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = lin(sd, p + ".proj_in", h)
    for k in range(depth):
        h = transformer_block(sd, f"{p}.transformer_blocks.{k}", h, ctx, heads)
    h = lin(sd, p + ".proj_out", h)
    return x + h.reshape(B, H, W, C).permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------------------------
# UNet2DConditionModel.forward (SDXL: addition_embed_type = "text_time")
# ------------------------------------------------------------------------------------------------------------------
def unet_forward(sd, cfg, sample, t, ctx, text_embeds, time_ids):
    sd = {k: v.float() for k, v in sd.items()}
    sample, ctx, text_embeds = sample.float(), ctx.float(), text_embeds.float()
    B = sample.shape[0]
    G = cfg["groups"]
    boc = cfg["block_out_channels"]
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1).expand(B)
    emb = lin(sd, "time_embedding.linear_2", F.silu(lin(sd, "time_embedding.linear_1", timestep_embedding(t, boc[0]))))
    tid = timestep_embedding(time_ids.reshape(-1), cfg["addition_time_embed_dim"]).reshape(B, -1)
    aug = torch.cat([text_embeds, tid], dim=-1)
    emb = emb + lin(sd, "add_embedding.linear_2", F.silu(lin(sd, "add_embedding.linear_1", aug)))

    h = conv(sd, "conv_in", sample)
    skips = [h]
    nb = len(boc)
    for i in range(nb):
        for j in range(cfg["layers_per_block"]):
            h = resnet(sd, f"down_blocks.{i}.resnets.{j}", h, emb, G, 1e-5)
            if cfg["down_attn"][i]:
                h = transformer2d(sd, f"down_blocks.{i}.attentions.{j}", h, ctx, cfg["heads"][i], cfg["transformer_layers"][i], G)
            skips.append(h)
        if i < nb - 1:
            h = conv(sd, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=1)
            skips.append(h)
    h = resnet(sd, "mid_block.resnets.0", h, emb, G, 1e-5)
    h = transformer2d(sd, "mid_block.attentions.0", h, ctx, cfg["heads"][-1], cfg["transformer_layers"][-1], G)
    h = resnet(sd, "mid_block.resnets.1", h, emb, G, 1e-5)
    for i in range(nb):
        r = nb - 1 - i  # mirrors down block r
        for j in range(cfg["layers_per_block"] + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(sd, f"up_blocks.{i}.resnets.{j}", h, emb, G, 1e-5)
            if cfg["down_attn"][r]:
                h = transformer2d(sd, f"up_blocks.{i}.attentions.{j}", h, ctx, cfg["heads"][r], cfg["transformer_layers"][r], G)
        if i < nb - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = conv(sd, f"up_blocks.{i}.upsamplers.0.conv", h)
    h = F.silu(gn(sd, "conv_norm_out", h, G, 1e-5))
    return conv(sd, "conv_out", h)


# ------------------------------------------------------------------------------------------------------------------
# AutoencoderKL
# ------------------------------------------------------------------------------------------------------------------
def vae_attention(sd, p, x, groups):
    """VAE mid-block Attention: single head over H*W tokens, GroupNorm in front, residual."""
    B, C, H, W = x.shape
    h = gn(sd, p + ".group_norm", x, groups, 1e-6).reshape(B, C, H * W).transpose(1, 2)
    q, k, v = lin(sd, p + ".to_q", h), lin(sd, p + ".to_k", h), lin(sd, p + ".to_v", h)
    o = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1) @ v
    o = lin(sd, p + ".to_out.0", o)
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def vae_decode(sd, cfg, z):
    """AutoencoderKL.decode(z) (z already divided by scaling_factor by the caller) -> image in [-1, 1]."""
    sd = {k: v.float() for k, v in sd.items()}
    G = cfg["groups"]
    rev = list(reversed(cfg["block_out_channels"]))
    h = conv(sd, "post_quant_conv", z.float(), padding=0)
    h = conv(sd, "decoder.conv_in", h)
    h = resnet(sd, "decoder.mid_block.resnets.0", h, None, G, 1e-6)
    h = vae_attention(sd, "decoder.mid_block.attentions.0", h, G)
    h = resnet(sd, "decoder.mid_block.resnets.1", h, None, G, 1e-6)
    for i in range(len(rev)):
        for j in range(cfg["layers_per_block"] + 1):
            h = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, None, G, 1e-6)
        if i < len(rev) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h)
    h = F.silu(gn(sd, "decoder.conv_norm_out", h, G, 1e-6))
    return conv(sd, "decoder.conv_out", h)


def vae_encode_mode(sd, cfg, x):
    """AutoencoderKL.encode(x).latent_dist.mode() = mean half of the moments."""
    sd = {k: v.float() for k, v in sd.items()}
    G = cfg["groups"]
    boc = cfg["block_out_channels"]
    h = conv(sd, "encoder.conv_in", x.float())
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"]):
            h = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, None, G, 1e-6)
        if i < len(boc) - 1:
            h = conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = resnet(sd, "encoder.mid_block.resnets.0", h, None, G, 1e-6)
    h = vae_attention(sd, "encoder.mid_block.attentions.0", h, G)
    h = resnet(sd, "encoder.mid_block.resnets.1", h, None, G, 1e-6)
    h = conv(sd, "encoder.conv_out", F.silu(gn(sd, "encoder.conv_norm_out", h, G, 1e-6)))
    moments = conv(sd, "quant_conv", h, padding=0)
    return moments[:, : cfg["latent_channels"]]


# ------------------------------------------------------------------------------------------------------------------
# EulerDiscreteScheduler (SDXL scheduler_config: scaled_linear betas, leading spacing, steps_offset 1, epsilon)
# ------------------------------------------------------------------------------------------------------------------
class Euler:
    def __init__(self, num_train=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float64) ** 2
        ac = torch.cumprod(1.0 - betas, dim=0)
        self.sigmas_all = ((1 - ac) / ac) ** 0.5
        self.num_train, self.steps_offset = num_train, steps_offset

    def set_timesteps(self, n):
        ratio = self.num_train // n
        ts = (torch.arange(n, dtype=torch.float64) * ratio).flip(0) + self.steps_offset
        idx = torch.arange(self.num_train, dtype=torch.float64)
        # np.interp(timesteps, arange(N), sigmas)
        lo = ts.floor().long().clamp(max=self.num_train - 1)
        hi = (lo + 1).clamp(max=self.num_train - 1)
        w = ts - lo.double()
        sig = self.sigmas_all[lo] * (1 - w) + self.sigmas_all[hi] * w
        self.timesteps = ts.float()
        self.sigmas = torch.cat([sig, torch.zeros(1, dtype=torch.float64)]).float()
        self.init_noise_sigma = float((self.sigmas.max() ** 2 + 1) ** 0.5)  # "leading" spacing
        del idx
        return self

    def scale_model_input(self, x, i):
        return x / ((self.sigmas[i] ** 2 + 1) ** 0.5)

    def step(self, eps, i, x):
        s, sn = self.sigmas[i], self.sigmas[i + 1]
        x0 = x - s * eps
        d = (x - x0) / s
        return x + d * (sn - s)


# ------------------------------------------------------------------------------------------------------------------
# pipelines
# ------------------------------------------------------------------------------------------------------------------
def add_time_ids(h, w):
    return torch.tensor([[h, w, 0, 0, h, w]], dtype=torch.float32)


def t2i_sample(unet_sd, unet_cfg, latents, prompt, pooled, neg_prompt, neg_pooled, steps=50, guidance=7.5, size=1024, return_all=False):
    """StableDiffusionXLPipeline.__call__ restated for given embeds (adapter_modules.py:156-167): 2-way CFG, batch order
    [negative, positive]; latents = standard normal, scaled by init_noise_sigma here."""
    sch = Euler().set_timesteps(steps)
    B = latents.shape[0]
    x = latents.float() * sch.init_noise_sigma
    ctx = torch.cat([neg_prompt, prompt]).float()
    txt = torch.cat([neg_pooled, pooled]).float()
    tid = add_time_ids(size, size).repeat(2 * B, 1)
    for i in range(steps):
        inp = sch.scale_model_input(torch.cat([x, x]), i)
        e = unet_forward(unet_sd, unet_cfg, inp, sch.timesteps[i], ctx, txt, tid)
        e_unc, e_txt = e.chunk(2)
        x = sch.step(e_unc + guidance * (e_txt - e_unc), i, x)
    return x


def edit_sample(unet_sd, unet_cfg, latents, image_latents, prompt, pooled, neg_prompt, neg_pooled, steps=50, guidance=7.5,
                image_guidance=1.5, size=1024):
    """StableDiffusionXLText2ImageAndEditPipeline.__call__ loop (pipeline_stable_diffusion_xl_t2i_edit.py:884-963):
    3-way CFG in order [text, image, uncond]; image latents [img, img, 0]; sigma-space combine (928-950)."""
    sch = Euler().set_timesteps(steps)
    B = latents.shape[0]
    x = latents.float() * sch.init_noise_sigma
    ctx = torch.cat([prompt, neg_prompt, neg_prompt]).float()
    txt = torch.cat([pooled, neg_pooled, neg_pooled]).float()
    tid = add_time_ids(size, size).repeat(3 * B, 1)
    img = torch.cat([image_latents, image_latents, torch.zeros_like(image_latents)]).float()
    for i in range(steps):
        raw = torch.cat([x] * 3)
        inp = torch.cat([sch.scale_model_input(raw, i), img], dim=1)
        e = unet_forward(unet_sd, unet_cfg, inp, sch.timesteps[i], ctx, txt, tid)
        s = sch.sigmas[i]
        e = raw - s * e                                   # to x0-space, using the UN-scaled latent input (:931)
        e_txt, e_img, e_unc = e.chunk(3)
        e = e_unc + guidance * (e_txt - e_img) + image_guidance * (e_img - e_unc)
        e = (e - x) / (-s)                                # back to eps (:949-950)
        x = sch.step(e, i, x)
    return x


def postprocess(img):
    """VaeImageProcessor.postprocess: [-1,1] -> uint8 HWC."""
    x = (img / 2 + 0.5).clamp(0, 1)
    return (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8)

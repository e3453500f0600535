"""De-tokenizer adapters end to end on the CPU double: `SDXLAdapter.generate(image_embeds=…)` and `SDXLAdapterWithLatentImage.generate(…,
latent_image=…)` run unchanged (ops modules swapped for tests/fake_ops.py) against the same composition done with the oracle pieces
(adapter_modules.py:96-169, 249-287): negative conditioning = avg-pooled ViT(zeros) cached once, resampler on [positive | negative] rows, branch
order of the 2-way and 3-way loops, source-image latents from the VAE encoder's mode (un-scaled), uint8 / PIL outputs."""
import numpy as np
import pytest
import torch

import fake_ops
from oracle import resampler_xl as orx
from oracle import sdxl as osd
from oracle import vit as ovit
from seedx_b200 import adapter as adapter_mod
from seedx_b200 import resampler_xl as rxl_mod
from seedx_b200 import sampler as sampler_mod
from seedx_b200 import sdxl as sdxl_mod
from seedx_b200 import synth
from seedx_b200 import vit as vit_mod

CPU = torch.device("cpu")
VCFG = dict(width=208, layers=2, heads=2, mlp_width=520, output_dim=256, n_queries=256, patch=14)
RCFG = dict(synth.TINY_RESAMPLER_XL, embedding_dim=256)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.fixture
def parts(monkeypatch):
    for mod in (adapter_mod, rxl_mod, sampler_mod, sdxl_mod, vit_mod):
        monkeypatch.setattr(mod, "ops", fake_ops)
    vit_sd, r_sd, v_sd = synth.vit_state_dict(**VCFG), synth.resampler_xl_state_dict(RCFG), synth.vae_state_dict(synth.TINY_VAE)
    vit = vit_mod.VisionTransformerWithAttnPool(image_size=448, patch_size=14, width=208, layers=2, heads=2, mlp_ratio=2.5, output_dim=256)
    vit.device = CPU
    vit.load_state_dict(vit_sd)
    rx = rxl_mod.ResamplerXLV2(normalize=False, **RCFG)
    rx.device = CPU
    rx.load_state_dict(r_sd)
    vae = sdxl_mod.AutoencoderKL(synth.TINY_VAE, device="cpu")
    vae.load_state_dict(v_sd)
    return dict(vit=vit, rx=rx, vae=vae, vit_sd=vit_sd, r_sd=r_sd, v_sd=v_sd)


def _adapter(cls, parts, ucfg, B, hw, branches):
    u_sd = synth.unet_state_dict(ucfg)
    unet = sdxl_mod.UNet2DConditionModel(ucfg, device="cpu")
    unet.load_state_dict(u_sd)
    ad = cls(unet=unet, resampler=parts["rx"], vit_down=True)
    ad.device = CPU
    ad.init_pipe(vae=parts["vae"], scheduler=sdxl_mod.EulerDiscreteScheduler(), visual_encoder=parts["vit"], image_transform=None)
    ad._loops[(B, hw * 8, hw * 8)] = sampler_mod.DenoiseLoop(unet, ad.scheduler, B, (hw, hw), branches, use_graph=False)   # no CUDA graphs on the CPU
    return ad, u_sd


def test_t2i_adapter_composition(parts):
    ucfg = dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160)
    B, hw, steps = 2, 8, 3
    ad, u_sd = _adapter(adapter_mod.SDXLAdapter, parts, ucfg, B, hw, 2)
    feats = synth.randn("cpu_adapter_feats", (B, 64, 256))
    noise = synth.randn("cpu_adapter_noise", (B, 4, hw, hw))
    lat = ad.generate(image_embeds=feats, num_inference_steps=steps, height=hw * 8, width=hw * 8, latents=noise, input_image_size=224, output_type="latent")
    neg = ovit.vit_down(ovit.vit_forward(parts["vit_sd"], torch.zeros(1, 3, 224, 224), 2))               # pooled ViT(zeros): 64 tokens
    prompt, pooled = orx.resampler_xl(parts["r_sd"], RCFG, torch.cat([feats, neg.expand(B, -1, -1)]))
    ref = osd.t2i_sample(u_sd, ucfg, noise, prompt[:B], pooled[:B], prompt[B:], pooled[B:], steps=steps, guidance=7.5, size=hw * 8)
    assert rel(lat, ref) < 1e-2
    assert list(ad._neg_cache) == [(224, True)]                                                       # computed once, re-used
    imgs = ad.generate(image_embeds=feats, num_inference_steps=steps, height=hw * 8, width=hw * 8, latents=noise, input_image_size=224)
    want = osd.postprocess(osd.vae_decode(parts["v_sd"], synth.TINY_VAE, ref / synth.TINY_VAE["scaling_factor"]))
    assert len(imgs) == B and imgs[0].size == (hw * 8, hw * 8) and len(ad._neg_cache) == 1
    assert np.abs(np.asarray(imgs[0]).astype(np.int32) - want[0].numpy().astype(np.int32)).mean() < 2.0


def test_edit_adapter_composition(parts):
    ucfg = dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160)
    B, hw, steps = 1, 8, 3
    ad, _ = _adapter(adapter_mod.SDXLAdapterWithLatentImage, parts, ucfg, B, hw, 3)
    assert ad.unet.cfg["in_channels"] == 8                              # widened at construction (adapter_modules.py:183-198); weights for 4..7 zero
    u_sd = synth.unet_state_dict(dict(ucfg, in_channels=8))
    ad.load_state_dict({"unet." + k: v for k, v in u_sd.items()})      # the full fine-tune checkpoint replaces the UNet
    ad._loops[(B, hw * 8, hw * 8)] = sampler_mod.DenoiseLoop(ad.unet, ad.scheduler, B, (hw, hw), 3, use_graph=False)
    feats = synth.randn("cpu_edit_feats", (B, 64, 256))
    noise = synth.randn("cpu_edit_noise2", (B, 4, hw, hw))
    src = synth.randn("cpu_edit_src", (B, 3, hw * 8, hw * 8)).clamp(-1, 1)
    lat = ad.generate(image_embeds=feats, latent_image=src, num_inference_steps=steps, height=hw * 8, width=hw * 8, latents=noise, input_image_size=224,
                      output_type="latent")
    neg = ovit.vit_down(ovit.vit_forward(parts["vit_sd"], torch.zeros(1, 3, 224, 224), 2))
    prompt, pooled = orx.resampler_xl(parts["r_sd"], RCFG, torch.cat([feats, neg.expand(B, -1, -1)]))
    il = osd.vae_encode_mode(parts["v_sd"], synth.TINY_VAE, src)                                       # latent_dist.mode(), NOT x 0.13025
    ref = osd.edit_sample(u_sd, dict(ucfg, in_channels=8), noise, il, prompt[:B], pooled[:B], prompt[B:], pooled[B:], steps=steps, size=hw * 8)
    assert rel(lat, ref) < 1e-2


def test_edit_adapter_matches_the_reference_adapter_and_pipeline(parts):
    """product SDXLAdapterWithLatentImage.generate (CPU double) vs the reference's OWN adapter + pipeline output (tests/golden/edit_adapter_tiny.pt)"""
    import os
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "edit_adapter_tiny.pt"))
    ucfg = dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160, in_channels=8)
    B, hw = 1, 8
    ad, _ = _adapter(adapter_mod.SDXLAdapterWithLatentImage, parts, ucfg, B, hw, 3)
    feats = synth.randn("edit_golden_feats", (B, 64, 256))
    noise = synth.randn("edit_golden_noise", (B, 4, hw, hw))
    src = synth.randn("edit_golden_src", (B, 3, hw * 8, hw * 8)).clamp(-1, 1)
    p, n, pp, npool = ad.get_image_embeds(image_embeds=feats, return_negative=True, image_size=224)
    assert rel(p, g["prompt"]) < 2e-3 and rel(n, g["neg_prompt"]) < 2e-3 and rel(pp, g["pooled"]) < 2e-3 and rel(npool, g["neg_pooled"]) < 2e-3
    tp, tn, tpp, tnp = ad.get_image_embeds(image_tensor=synth.image("edit_golden_img224", 1, 224), return_negative=True)     # un-pooled 256-token path
    assert rel(tp, g["tensor_prompt"]) < 2e-3 and rel(tn, g["tensor_neg_prompt"]) < 2e-3
    assert rel(tpp, g["tensor_pooled"]) < 2e-3 and rel(tnp, g["tensor_neg_pooled"]) < 2e-3
    lat = ad.generate(image_embeds=feats, latent_image=src, num_inference_steps=g["steps"], height=hw * 8, width=hw * 8, latents=noise, input_image_size=224,
                      guidance_scale=7.5, image_guidance_scale=1.5, output_type="latent")
    assert rel(lat, g["latents"]) < 1e-2
    u8 = ad.generate(image_embeds=feats, latent_image=src, num_inference_steps=g["steps"], height=hw * 8, width=hw * 8, latents=noise, input_image_size=224,
                     guidance_scale=7.5, image_guidance_scale=1.5, output_type="uint8")
    want = ((g["image"] / 2 + 0.5).clamp(0, 1) * 255).round().permute(0, 2, 3, 1)
    assert (u8.float() - want).abs().mean() < 2.0

"""Stage-3 host wiring on the CPU double: `seedx_b200.sdxl.UNet2DConditionModel` / `AutoencoderKL` run unchanged over tests/fake_ops.py and are
compared with oracle/sdxl.py (tiny configs): weight packing (padded-channel conv layout, fused QKV / KV, interleaved GEGLU rows, dense stride-2
convs), skip-connection order and channel concat, hoisted cross-attention K/V, the added time-id embedding, 4- and 8-channel conv_in, VAE
mid attention through the transposed-V products, asymmetric stride-2 padding of the VAE encoder, uint8 post-processing."""
import pytest
import torch

import fake_ops
from oracle import sdxl as osd
from seedx_b200 import sdxl as sdxl_mod
from seedx_b200 import synth


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.fixture
def patched(monkeypatch):
    monkeypatch.setattr(sdxl_mod, "ops", fake_ops)


@pytest.mark.parametrize("in_ch", [4, 8])
def test_unet_host_wiring_vs_oracle(patched, in_ch):
    cfg = dict(synth.TINY_UNET, in_channels=in_ch)
    sd = synth.unet_state_dict(cfg)
    B, hw = 2, 16
    x = synth.randn("cpu_unet_x", (B, in_ch, hw, hw))
    ctx = synth.randn("cpu_unet_ctx", (B, 16, cfg["cross_attention_dim"]))
    te = synth.randn("cpu_unet_te", (B, cfg["text_embed_dim"]))
    tid = torch.tensor([[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]]).repeat(B, 1)
    ref = osd.unet_forward(sd, cfg, x, 601.0, ctx, te, tid)
    m = sdxl_mod.UNet2DConditionModel(cfg, device="cpu")
    m.load_state_dict(sd)
    assert m.cfg["in_channels"] == in_ch
    out = m(x, 601.0, ctx, added_cond_kwargs=dict(text_embeds=te, time_ids=tid))
    assert out.shape == ref.shape and rel(out, ref) < 5e-3


def test_partial_adapter_checkpoint_updates_the_unet_in_place(patched):
    """adapter checkpoints saved with full_ft=False hold only the UNet's `to_k` / `to_v` (+ the widened `conv_in` of the edit variant),
    loaded with strict=False in the reference (adapter_modules.py:20-33,59-65): they merge into the loaded base UNet, the packed device tensors
    keep their storage (captured graphs / cached loops stay valid), every other parameter keeps its base value, unknown keys are reported."""
    from seedx_b200.adapter import SDXLAdapterWithLatentImage
    cfg = dict(synth.TINY_UNET, in_channels=4)
    base = synth.unet_state_dict(cfg)
    m = sdxl_mod.UNet2DConditionModel(cfg, device="cpu")
    m.load_state_dict({k: v.clone() for k, v in base.items()})
    ptrs = {k: v[0].data_ptr() for k, v in m._reg.slots.items()}
    assert set(m._reg.slots) == set(base)                             # every checkpoint key has a packed destination
    part = {k: synth.randn("ft:" + k, v.shape, v.shape[-1] ** -0.5) for k, v in base.items() if k.endswith(("to_k.weight", "to_v.weight"))}
    assert len(part) >= 8
    part["conv_in.weight"] = torch.cat([base["conv_in.weight"], synth.randn("ft:conv_in", base["conv_in.weight"].shape, 0.05)], dim=1)
    ad = SDXLAdapterWithLatentImage.__new__(SDXLAdapterWithLatentImage)
    ad.unet, ad.resampler = m, None
    ad.load_state_dict({**{"unet." + k: v for k, v in part.items()}, "unet.not_a_parameter": torch.zeros(1)})
    assert m.cfg["in_channels"] == 8
    assert ptrs == {k: v[0].data_ptr() for k, v in m._reg.slots.items()}
    merged = {**base, **part}
    cfg8 = dict(cfg, in_channels=8)
    B, hw = 1, 16
    x = synth.randn("cpu_unet_x8", (B, 8, hw, hw))
    ctx = synth.randn("cpu_unet_ctx", (B, 16, cfg["cross_attention_dim"]))
    te = synth.randn("cpu_unet_te", (B, cfg["text_embed_dim"]))
    tid = torch.tensor([[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]])
    ref = osd.unet_forward(merged, cfg8, x, 301.0, ctx, te, tid)
    out = m(x, 301.0, ctx, added_cond_kwargs=dict(text_embeds=te, time_ids=tid))
    assert rel(out, ref) < 5e-3
    assert rel(out, osd.unet_forward({**base, "conv_in.weight": part["conv_in.weight"]}, cfg8, x, 301.0, ctx, te, tid)) > 5e-2   # the update took effect
    missing, unexpected = m.load_state_dict({"mid_block.resnets.0.conv1.bias": base["mid_block.resnets.0.conv1.bias"], "bogus": torch.zeros(1)})
    assert unexpected == ["bogus"] and len(missing) == len(base) - 1
    with pytest.raises(sdxl_mod.SeedxError):
        m.load_state_dict({"conv_out.weight": torch.zeros(5, 7, 3, 3)})
    # LayerNorms are folded into the projections that read them: a checkpoint that changes one must bring those projections along
    blk = "down_blocks.1.attentions.0.transformer_blocks.0."
    g2 = {blk + "norm1.weight": synth.randn("ft:g", base[blk + "norm1.weight"].shape, 0.2, 1.0), blk + "norm1.bias": synth.randn("ft:b", base[blk + "norm1.bias"].shape, 0.2)}
    with pytest.raises(sdxl_mod.SeedxError):
        m.load_state_dict(dict(g2))
    qkv = {blk + f"attn1.{n}.weight": merged[blk + f"attn1.{n}.weight"] for n in ("to_q", "to_k", "to_v")}
    m.load_state_dict({**g2, **qkv})
    merged.update(g2)
    assert rel(m(x, 301.0, ctx, added_cond_kwargs=dict(text_embeds=te, time_ids=tid)), osd.unet_forward(merged, cfg8, x, 301.0, ctx, te, tid)) < 5e-3


def test_vae_host_wiring_vs_oracle(patched):
    cfg = synth.TINY_VAE
    sd = synth.vae_state_dict(cfg)
    m = sdxl_mod.AutoencoderKL(cfg, device="cpu")
    m.load_state_dict(sd)
    z = synth.randn("cpu_vae_z", (1, 4, 8, 8))
    ref = osd.vae_decode(sd, cfg, z / cfg["scaling_factor"])
    img = m.decode(z, scale=1.0 / cfg["scaling_factor"])
    assert img.shape == ref.shape == (1, 3, 64, 64) and rel(img, ref) < 5e-3
    u8 = fake_ops.image_to_u8(m.decode_nhwc(z, scale=1.0 / cfg["scaling_factor"]))
    assert (u8.int() - osd.postprocess(ref).int()).abs().float().mean() < 1.0
    x = synth.randn("cpu_vae_img", (1, 3, 64, 64)).clamp(-1, 1)
    lat = m.encode_mode(x)
    assert lat.shape == (1, 4, 8, 8) and rel(lat, osd.vae_encode_mode(sd, cfg, x)) < 5e-3


def _overflowing_vae(cfg):
    """synthetic VAE whose un-normalised activations leave the fp16 range (like the stock SDXL VAE): the layers that write the residual stream
    are scaled up; every GroupNorm input is then O(1e4..1e6)"""
    sd = synth.vae_state_dict(cfg)
    hot = ("conv2.weight", "conv2.bias", "conv_shortcut.weight", "conv_shortcut.bias", "conv_in.weight", "conv_in.bias")
    return {k: (v * 300 if k.endswith(hot) and "quant" not in k else v) for k, v in sd.items()}


def test_vae_force_upcast_keeps_the_stream_in_fp16_range(patched, tmp_path):
    """`force_upcast` (the stock SDXL VAE config; the reference then runs the VAE in fp32, pipeline...edit.py:569-586,965-975): the plain fp16
    stream overflows on such weights, the scaled stream (stream_scale 2^-7, exact in fp32) matches the fp32 oracle; config.json is honoured."""
    import json
    cfg = dict(synth.TINY_VAE)
    big = _overflowing_vae(cfg)
    z = synth.randn("cpu_vae_z", (1, 4, 8, 8))
    x = synth.randn("cpu_vae_img", (1, 3, 64, 64)).clamp(-1, 1)
    ref_img, ref_lat = osd.vae_decode(big, cfg, z / cfg["scaling_factor"]), osd.vae_encode_mode(big, cfg, x)
    plain = sdxl_mod.AutoencoderKL(dict(cfg, force_upcast=False), device="cpu")
    plain.load_state_dict(big)
    assert plain.stream_scale == 1.0 and not torch.isfinite(plain.decode(z, scale=1.0 / cfg["scaling_factor"])).all()     # what the upcast is for
    json.dump(dict(block_out_channels=list(cfg["block_out_channels"]), layers_per_block=cfg["layers_per_block"], latent_channels=4,
                   scaling_factor=cfg["scaling_factor"], force_upcast=True), open(tmp_path / "config.json", "w"))
    torch.save(big, tmp_path / "diffusion_pytorch_model.bin")
    m = sdxl_mod.AutoencoderKL.from_pretrained(str(tmp_path), device="cpu")
    assert m.cfg["force_upcast"] is True and m.stream_scale == 2.0 ** -7 and m.config.force_upcast is False
    img, lat = m.decode(z, scale=1.0 / cfg["scaling_factor"]), m.encode_mode(x)
    assert torch.isfinite(img).all() and rel(img, ref_img) < 5e-3 and rel(lat, ref_lat) < 5e-3
    # on weights that do not overflow both modes agree with the oracle equally well
    sd = synth.vae_state_dict(cfg)
    ref = osd.vae_decode(sd, cfg, z / cfg["scaling_factor"])
    for fu in (False, True):
        v = sdxl_mod.AutoencoderKL(dict(cfg, force_upcast=fu), device="cpu")
        v.load_state_dict(sd)
        assert rel(v.decode(z, scale=1.0 / cfg["scaling_factor"]), ref) < 5e-3

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

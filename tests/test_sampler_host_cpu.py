"""Host logic of the de-tokenizer sampling loops on the CPU: `seedx_b200.sampler.DenoiseLoop` runs unchanged over tests/fake_ops.py and a UNet
stand-in that evaluates the oracle UNet (oracle/sdxl.py).  Checked: the Euler schedule it walks, `scale_model_input`, the branch order and
CFG arithmetic of the t2i ([negative, positive]) and edit ([text, image, uncond], sigma space, un-scaled image latents) loops, re-use of the
static conditioning buffers across requests — against oracle.t2i_sample / oracle.edit_sample (which restate the reference pipelines)."""
import pytest
import torch

import fake_ops
from oracle import sdxl as osd
from seedx_b200 import sampler as sampler_mod
from seedx_b200 import synth
from seedx_b200.sdxl import EulerDiscreteScheduler


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


class OracleUNet:
    """forward_nhwc / prepare_cond surface of seedx_b200.sdxl.UNet2DConditionModel, arithmetic by the oracle"""

    def __init__(self, sd, cfg):
        self.sd, self.cfg, self.device = sd, cfg, torch.device("cpu")
        self.calls = 0

    def prepare_cond(self, ctx, text_embeds, time_ids):
        return dict(ctx=ctx.clone().float(), te=text_embeds.clone().float(), tid=time_ids.clone().float())

    def forward_nhwc(self, x_in, t_dev, cond):
        self.calls += 1
        assert float(t_dev.min()) == float(t_dev.max())
        x = x_in[..., : self.cfg["in_channels"]].float().permute(0, 3, 1, 2).contiguous()
        eps = osd.unet_forward(self.sd, self.cfg, x, float(t_dev[0]), cond["ctx"], cond["te"], cond["tid"])
        return eps.permute(0, 2, 3, 1).contiguous().float()


def _cond(cfg, B, tag):
    T = 16
    return (synth.randn(tag + "p", (B, T, cfg["cross_attention_dim"])), synth.randn(tag + "pp", (B, cfg["text_embed_dim"])),
            synth.randn(tag + "n", (B, T, cfg["cross_attention_dim"])), synth.randn(tag + "np", (B, cfg["text_embed_dim"])))


@pytest.fixture
def patched(monkeypatch):
    monkeypatch.setattr(sampler_mod, "ops", fake_ops)


def test_t2i_loop_host_logic(patched):
    cfg = synth.TINY_UNET
    sd = synth.unet_state_dict(cfg)
    B, hw, steps = 2, 16, 4
    noise = synth.randn("cpu_t2i_noise", (B, 4, hw, hw))
    p, pp, n, npool = _cond(cfg, B, "cpu_t2i")
    ref = osd.t2i_sample(sd, cfg, noise, p, pp, n, npool, steps=steps, guidance=7.5, size=1024)
    unet = OracleUNet(sd, cfg)
    loop = sampler_mod.DenoiseLoop(unet, EulerDiscreteScheduler(), B, (hw, hw), 2, use_graph=False)
    tid = torch.tensor([[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]]).repeat(2 * B, 1)
    loop.set_condition(torch.cat([n, p]), torch.cat([npool, pp]), tid)
    lat = loop.run(noise, steps=steps, guidance=7.5).clone()
    assert unet.calls == steps and rel(lat, ref) < 2e-3                    # fp16 rounding of the UNet input only
    # a second request re-uses the same conditioning buffers (what keeps a captured graph valid) and still gets its own result
    cond_before = loop.cond
    p2, pp2, n2, npool2 = _cond(cfg, B, "cpu_t2i_second")
    noise2 = synth.randn("cpu_t2i_noise2", (B, 4, hw, hw))
    loop.set_condition(torch.cat([n2, p2]), torch.cat([npool2, pp2]), tid)
    assert loop.cond is cond_before
    ref2 = osd.t2i_sample(sd, cfg, noise2, p2, pp2, n2, npool2, steps=steps, guidance=7.5, size=1024)
    assert rel(loop.run(noise2, steps=steps, guidance=7.5), ref2) < 2e-3


def test_edit_loop_host_logic(patched):
    cfg = dict(synth.TINY_UNET, in_channels=8)
    sd = synth.unet_state_dict(cfg)
    B, hw, steps = 1, 16, 3
    noise = synth.randn("cpu_edit_noise", (B, 4, hw, hw))
    il = synth.randn("cpu_edit_il", (B, 4, hw, hw), 0.7)
    p, pp, n, npool = _cond(cfg, B, "cpu_edit")
    ref = osd.edit_sample(sd, cfg, noise, il, p, pp, n, npool, steps=steps)
    unet = OracleUNet(sd, cfg)
    loop = sampler_mod.DenoiseLoop(unet, EulerDiscreteScheduler(), B, (hw, hw), 3, use_graph=False)
    tid = torch.tensor([[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]]).repeat(3 * B, 1)
    loop.set_condition(torch.cat([p, n, n]), torch.cat([pp, npool, npool]), tid, image_latents=il)       # [text, image, uncond]
    lat = loop.run(noise, steps=steps, guidance=7.5, image_guidance=1.5)
    assert rel(lat, ref) < 2e-3
    # the uncond branch sees zero image latents, the other two the un-scaled source latents (pipeline...edit.py:544-546, 523)
    assert float(loop.unet_in[2 * B:, ..., 4:].abs().max()) == 0.0
    assert rel(loop.unet_in[:B, ..., 4:8].float().permute(0, 3, 1, 2), il) < 1e-3
    from seedx_b200._lib import SeedxError
    with pytest.raises(SeedxError):
        sampler_mod.DenoiseLoop(unet, EulerDiscreteScheduler(), B, (hw, hw), 4)


def test_scheduler_from_pretrained_refuses_configs_it_does_not_implement(tmp_path):
    """the SDXL-base scheduler_config.json loads; a config that asks for another spacing / prediction type / Karras sigmas would sample with the
    wrong sigmas without any error, so from_pretrained raises instead of silently assuming the SDXL defaults"""
    import json
    from seedx_b200._lib import SeedxError
    base = dict(_class_name="EulerDiscreteScheduler", beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, clip_sample=False,
                interpolation_type="linear", num_train_timesteps=1000, prediction_type="epsilon", sample_max_value=1.0, set_alpha_to_one=False,
                skip_prk_steps=True, steps_offset=1, timestep_spacing="leading", trained_betas=None, use_karras_sigmas=False)
    d = tmp_path / "scheduler"
    d.mkdir()
    json.dump(base, open(d / "scheduler_config.json", "w"))
    s = EulerDiscreteScheduler.from_pretrained(str(tmp_path), subfolder="scheduler").set_timesteps(50)
    o = osd.Euler().set_timesteps(50)
    assert torch.allclose(torch.tensor(s.sigmas), o.sigmas, rtol=1e-6) and abs(s.init_noise_sigma - o.init_noise_sigma) < 1e-5
    for bad in (dict(timestep_spacing="trailing"), dict(prediction_type="v_prediction"), dict(use_karras_sigmas=True), dict(beta_schedule="linear"),
                dict(trained_betas=[0.1, 0.2])):
        json.dump({**base, **bad}, open(d / "scheduler_config.json", "w"))
        with pytest.raises(SeedxError):
            EulerDiscreteScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")

"""De-tokenizer front (`seedx_b200.resampler_xl.ResamplerXLV2`) host wiring on the CPU double vs the REFERENCE module's own outputs
(tests/golden/resampler_xl.pt): perceiver layers over cat(LN(x), LN(latents)), fused unet_proj_1|2 head, AttentionPool2d on the mean token;
tiny and full-size (61 M parameters) configs, 64- and 256-token inputs."""
import os

import pytest
import torch

import fake_ops
from seedx_b200 import resampler_xl as rxl_mod
from seedx_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_resampler_xl_host_wiring_matches_reference_golden(monkeypatch, name):
    monkeypatch.setattr(rxl_mod, "ops", fake_ops)
    g = torch.load(os.path.join(GOLD, "resampler_xl.pt"))
    cfg = synth.TINY_RESAMPLER_XL if name == "tiny" else synth.RESAMPLER_XL
    m = rxl_mod.ResamplerXLV2(normalize=False, **cfg)
    m.device = torch.device("cpu")
    m.load_state_dict(synth.resampler_xl_state_dict(cfg))
    for n_tok in (64, 256):
        x = synth.randn(f"rxl_{name}_{n_tok}", (2, n_tok, cfg["embedding_dim"]))
        prompt, pooled = m(x)
        assert rel(prompt, g[f"{name}_{n_tok}_prompt"]) < 2e-3 and rel(pooled, g[f"{name}_{n_tok}_pooled"]) < 2e-3
    one = m(x[:1])                                       # batch 1 must not update the learned latents in place (round-1 aliasing bug)
    again = m(x[:1])
    assert torch.equal(one[0], again[0]) and rel(one[0], prompt[:1]) < 2e-3

"""Pin the CPU oracle (oracle/*.py) against golden vectors produced by the reference modules themselves
(tests/golden/make_golden.py).  fp32 vs fp32 on CPU: tolerance 2e-5 relative Frobenius (summation order only)."""
import os

import torch

from seedx_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_sincos_table_matches_reference():
    g = torch.load(os.path.join(GOLD, "vit_small.pt"))
    assert torch.allclose(synth.sincos_2d(256, 16), g["attn_pool_pos_embed"], atol=1e-6)


def test_vit_oracle_matches_reference():
    from oracle import vit
    g = torch.load(os.path.join(GOLD, "vit_small.pt"))
    cfg = g["cfg"]
    sd = synth.vit_state_dict(**cfg)
    for size in (448, 224):
        x = synth.image(f"vit_small_in_{size}", 2, size)
        out = vit.vit_forward(sd, x, cfg["heads"])
        assert out.shape == g[f"out_{size}"].shape
        assert rel(out, g[f"out_{size}"]) < 2e-5, size


def test_resampler_oracle_matches_reference():
    from oracle import vit
    g = torch.load(os.path.join(GOLD, "resamplers.pt"))
    for name, grid, E, heads, kv_dim, nkv in (("in", 8, 320, 2, 256, 256), ("out", 8, 256, 2, 320, 64)):
        sd = synth.resampler_state_dict(f"res_{name}.", grid, E, kv_dim)
        x = synth.randn(f"res_{name}_x", (3, nkv, kv_dim))
        out = vit.resampler(sd, f"res_{name}.", x, heads, 1e-5)
        assert rel(out, g[name]) < 2e-5, name

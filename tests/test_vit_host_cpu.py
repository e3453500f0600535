"""Stage-1 host wiring on the CPU: `seedx_b200.vit.VisionTransformerWithAttnPool` runs unchanged over tests/fake_ops.py (test double of the C
entry points) and is compared with the output of the REFERENCE module itself (tests/golden/vit_small.pt): weight packing (padded patch-embed K,
interleaved per-head QKV read in place, transposed proj), bicubic position tables for 448^2 and the native 224^2 grid, attention pooling."""
import os

import torch

import fake_ops
from seedx_b200 import synth
from seedx_b200 import vit as vit_mod

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_vit_host_wiring_matches_reference_golden(monkeypatch):
    monkeypatch.setattr(vit_mod, "ops", fake_ops)
    g = torch.load(os.path.join(GOLD, "vit_small.pt"))
    cfg = g["cfg"]
    m = vit_mod.VisionTransformerWithAttnPool(image_size=448, patch_size=14, width=cfg["width"], layers=cfg["layers"], heads=cfg["heads"],
                                              mlp_ratio=cfg["mlp_width"] / cfg["width"], n_queries=256, output_dim=cfg["output_dim"])
    m.device = torch.device("cpu")                      # the product pins cuda; the double has no device
    m.load_state_dict(synth.vit_state_dict(**cfg))
    m.to(dtype=torch.float32)
    for size in (448, 224):
        x = synth.image(f"vit_small_in_{size}", 2, size)
        out = m(x)
        assert out.shape == g[f"out_{size}"].shape
        assert rel(out, g[f"out_{size}"]) < 1e-3, size
    out2 = m(torch.cat([x, x]))
    assert torch.equal(out2[:2], out2[2:])

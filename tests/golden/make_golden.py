"""Generate golden vectors by running the UNMODIFIED reference modules (imported from /root/reference) on seeded synthetic
weights/inputs.  Run in the build container only:  python tests/golden/make_golden.py
Outputs small .pt fixtures next to this file; weights are NOT stored (they are a pure function of their names, see
seedx_b200/synth.py), only the reference outputs are.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from _ref_import import ref_module  # noqa: E402
from seedx_b200 import synth  # noqa: E402

VIT_SMALL = dict(width=208, layers=2, heads=2, mlp_width=520, output_dim=256, n_queries=256, patch=14)


def golden_vit():
    qv = ref_module("src.models.tokenizer.qwen_visual")
    cfg = VIT_SMALL
    torch.manual_seed(0)
    model = qv.VisionTransformerWithAttnPool(image_size=448, patch_size=cfg["patch"], width=cfg["width"], layers=cfg["layers"],
                                             heads=cfg["heads"], mlp_ratio=cfg["mlp_width"] / cfg["width"], n_queries=cfg["n_queries"],
                                             output_dim=cfg["output_dim"]).eval()
    sd = synth.vit_state_dict(**cfg)
    ref_pos = model.attn_pool.pos_embed.detach().clone()          # the reference's own sincos table
    sd_no_pos = {k: v for k, v in sd.items() if k != "attn_pool.pos_embed"}
    missing, unexpected = model.load_state_dict(sd_no_pos, strict=False)
    assert list(missing) == ["attn_pool.pos_embed"] and not unexpected, (missing, unexpected)
    out = {}
    with torch.no_grad():
        for size in (448, 224):
            x = synth.image(f"vit_small_in_{size}", 2, size)
            out[f"out_{size}"] = model(x).float()
    out["attn_pool_pos_embed"] = ref_pos
    out["cfg"] = cfg
    torch.save(out, os.path.join(HERE, "vit_small.pt"))
    print("vit_small", {k: tuple(v.shape) for k, v in out.items() if hasattr(v, "shape")})


def golden_resamplers():
    qv = ref_module("src.models.tokenizer.qwen_visual")
    out = {}
    # (name, grid, embed_dim, heads, kv_dim, n_kv)  -- scaled-down input_resampler (d=160, 16x16 kv grid) / output_resampler
    for name, grid, E, heads, kv_dim, nkv in (("in", 8, 320, 2, 256, 256), ("out", 8, 256, 2, 320, 64)):
        m = qv.Resampler(grid_size=grid, embed_dim=E, num_heads=heads, kv_dim=kv_dim).eval()
        sd = synth.resampler_state_dict(f"res_{name}.", grid, E, kv_dim)
        sd = {k[len(f"res_{name}."):]: v for k, v in sd.items()}
        m.load_state_dict(sd, strict=True)
        x = synth.randn(f"res_{name}_x", (3, nkv, kv_dim))
        with torch.no_grad():
            out[name] = m(x).float()
    torch.save(out, os.path.join(HERE, "resamplers.pt"))
    print("resamplers", {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["vit", "resamplers"]
    if "vit" in which:
        golden_vit()
    if "resamplers" in which:
        golden_resamplers()

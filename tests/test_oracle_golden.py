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


def test_llama_oracle_matches_reference():
    """forward (prefill logits / hidden), cached decode, logits processor and the greedy loop vs the reference's own outputs."""
    from oracle import llm
    g = torch.load(os.path.join(GOLD, "llama_tiny.pt"))
    cfg = synth.TINY_LLAMA
    sd = synth.llama_state_dict(cfg)
    tok = synth.SynthTokenizer(vocab=cfg["vocab"])
    logits, hid, cache = llm.llama_forward(sd, cfg, g["embeds"], 0, None)
    assert rel(logits, g["prefill_logits"]) < 2e-5 and rel(hid, g["prefill_hidden"]) < 2e-5
    img_ids = tok.encode("".join(["<img>"] + ["<img_{:05d}>".format(i) for i in range(64)] + ["</img>"]))
    # processor restatement
    assert torch.equal(llm.image_token_processor(img_ids, 7, g["proc_in"][0]), g["proc_out_text"][0])
    assert torch.equal(llm.image_token_processor(img_ids, tok.encode("<img_00010>")[0], g["proc_in"][0]), g["proc_out_img"][0])
    # greedy loops
    gen, h = llm.greedy_generate(sd, cfg, g["ids"], g["embeds"], img_ids, 16)
    assert gen == g["text_gen_ids"] and rel(h, g["text_hidden"]) < 2e-5
    ids_b = g["ids"] + [tok.encode("<img>")[0]]
    emb_b = torch.cat([g["embeds"], sd["model.embed_tokens.weight"][ids_b[-1]][None]])
    gen, h = llm.greedy_generate(sd, cfg, ids_b, emb_b, img_ids, 72)
    assert gen == g["img_gen_ids"] and rel(h, g["img_hidden"]) < 2e-5


def test_resampler_xl_oracle_matches_reference():
    from oracle import resampler_xl as orx
    g = torch.load(os.path.join(GOLD, "resampler_xl.pt"))
    for name, cfg in (("tiny", synth.TINY_RESAMPLER_XL), ("full", synth.RESAMPLER_XL)):
        sd = synth.resampler_xl_state_dict(cfg)
        for n_tok in (64, 256):
            x = synth.randn(f"rxl_{name}_{n_tok}", (2, n_tok, cfg["embedding_dim"]))
            p, pooled = orx.resampler_xl(sd, cfg, x)
            assert rel(p, g[f"{name}_{n_tok}_prompt"]) < 5e-4          # golden prompt stored as fp16
            assert rel(pooled, g[f"{name}_{n_tok}_pooled"]) < 2e-5


def test_host_preprocessing_matches_reference():
    """seedx_b200.preprocess (get_transform, process_anyres_image) vs the reference functions' outputs on seeded noise images."""
    import numpy as np
    from PIL import Image
    from seedx_b200 import preprocess as pp
    g = torch.load(os.path.join(GOLD, "preprocess.pt"))
    rng = np.random.RandomState(0)
    for c in g["cases"]:
        w, h = c["size"]
        img = Image.fromarray(rng.randint(0, 255, (h, w, 3), dtype=np.uint8))
        views, pos = pp.process_anyres_image(img, pp.get_transform("clip", keep_ratio=False, image_size=448), g["grids"], 448)
        assert tuple(views.shape) == c["shape"] and torch.equal(pos, c["pos"])
        assert abs(float(views.double().sum()) - c["sum"]) < 1e-6 * max(1.0, abs(c["abssum"]))
        assert torch.equal(views[0, :, :4, :4], c["first"]) and torch.equal(views[-1, :, -4:, -4:], c["last"])
        tk = pp.get_transform("clip", keep_ratio=True, image_size=448)(img)
        assert abs(float(tk.double().sum()) - c["keep_sum"]) < 1e-6 * max(1.0, abs(c["abssum"]))

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


def test_llama_lora_oracle_and_merge_match_reference_peft():
    """un-merged LoRA forward restated in the oracle, the load-time merge of seedx_b200.lora and the vocabulary growth, all against the
    reference's get_peft_model_with_resize_embedding run over its vendored PEFT 0.4.0 (golden = un-merged fp32 logits)."""
    from oracle import llm
    from seedx_b200 import lora
    g = torch.load(os.path.join(GOLD, "llama_lora_tiny.pt"))
    cfg = dict(synth.TINY_LLAMA)
    base = llm.resize_embeddings(synth.llama_state_dict(cfg), g["new_vocab"])
    assert torch.allclose(base["model.embed_tokens.weight"][cfg["vocab"]:], g["embed_new_rows"], atol=1e-6)
    assert torch.allclose(base["lm_head.weight"][cfg["vocab"]:], g["head_new_rows"], atol=1e-6)
    cfg["vocab"] = g["new_vocab"]
    ft = synth.lora_fixture(g["shapes"])                                  # PEFT key names exactly as the reference model reports them
    scaling = g["lora_alpha"] / g["r"]
    # (1) oracle, un-merged
    _, pairs, saved = lora.split_peft_state_dict(ft)
    assert len(pairs) == 7 * cfg["layers"] and len(saved) == 2 * cfg["layers"] + 1
    sd_u = {**base, **saved}
    lsd = {}
    for mod, (a, b) in pairs.items():
        lsd[mod + ".lora_A.weight"], lsd[mod + ".lora_B.weight"] = a, b
    x = sd_u["model.embed_tokens.weight"][torch.tensor(g["ids"])]
    logits, hid, _ = llm.llama_forward(sd_u, cfg, x, 0, None, lora=dict(scaling=scaling, sd=lsd))
    assert rel(logits, g["logits"]) < 2e-5 and rel(hid, g["hidden"]) < 2e-5
    # (2) product merge: a PEFT-named state dict holding base + adapters -> plain HF names
    full = {"base_model.model." + k: v for k, v in base.items() if "norm" not in k and "rotary_emb" not in k}   # PEFT wraps the saved norms: no plain key for them
    full.update(ft)
    assert set(full) == {k for k in g["peft_keys"] if ".original_module." not in k and "rotary_emb" not in k}
    merged = lora.merge_lora_state_dict(full, scaling)
    assert set(merged) == {k for k in base if "rotary_emb" not in k}
    logits_m, hid_m, _ = llm.llama_forward(merged, cfg, x, 0, None)
    assert rel(logits_m, g["logits"]) < 2e-5 and rel(hid_m, g["hidden"]) < 2e-5
    # the adapters matter: without them the logits are far away
    logits_0, _, _ = llm.llama_forward(base, cfg, x, 0, None)
    assert rel(logits_0, g["logits"]) > 0.05


def test_product_logits_processor_class_matches_reference():
    """seedx_b200.llm.AutoImageTokenGenerationProcessor (host form of the rule the device kernel applies) vs the reference class's outputs."""
    from seedx_b200.llm import AutoImageTokenGenerationProcessor
    g = torch.load(os.path.join(GOLD, "llama_tiny.pt"))
    tok = synth.SynthTokenizer(vocab=synth.TINY_LLAMA["vocab"])
    proc = AutoImageTokenGenerationProcessor(tok, num_img_gen_tokens=64)
    assert len(proc.img_ids_list) == 66
    assert torch.equal(proc(torch.tensor([[5, 6, 7]]), g["proc_in"].clone()), g["proc_out_text"])
    assert torch.equal(proc(torch.tensor([[5, tok.encode("<img_00010>")[0]]]), g["proc_in"].clone()), g["proc_out_img"])


def test_lvlm_generate_oracle_matches_the_reference_generate():
    """oracle/llm.py::lvlm_generate vs the reference's OWN ContinuousLVLM.generate (tests/golden/agent_tiny.pt: only `llm.generate` was substituted by
    the restated greedy loop): generated text, image-span detection and the output-resampler features of a 2-image, 3-turn prompt"""
    from oracle import llm
    g = torch.load(os.path.join(GOLD, "agent_tiny.pt"))
    cfg = synth.TINY_LLAMA
    tok = synth.SynthTokenizer(vocab=cfg["vocab"])
    image_embeds = synth.randn("agent_golden_img", (3, 256, 320))
    out = llm.lvlm_generate(synth.llama_state_dict(cfg), synth.agent_state_dict(cfg["hidden"], 320), cfg, tok, g["input_ids"][0].tolist(), image_embeds,
                            g["ids_cmp_mask"][0], torch.ones((3, 64), dtype=torch.bool), g["patch_pos"], 70, eos_id=2)
    assert out["text"] == g["text"] and out["has_img_output"] == g["has_img_output"] and out["num_gen_imgs"] == g["num_gen_imgs"] == 1
    assert rel(out["img_gen_feat"], g["img_gen_feat"]) < 2e-5


def test_edit_pipeline_oracle_matches_the_reference_adapter_and_pipeline():
    """oracle composition (ViT(zeros) pooled negative -> ResamplerXLV2 -> VAE-encoded source latents -> oracle.edit_sample) vs the output of the
    reference's OWN SDXLAdapterWithLatentImage.generate + StableDiffusionXLText2ImageAndEditPipeline.__call__ (tests/golden/edit_adapter_tiny.pt; the
    UNet / VAE / scheduler objects handed to the reference code were backed by this same oracle, so what is pinned is everything around them: conditioning,
    branch order [text, image, uncond], time ids, un-scaled image latents, sigma-space 3-way CFG, the Euler walk)"""
    from oracle import resampler_xl as orx, sdxl as osd, vit as ovit
    g = torch.load(os.path.join(GOLD, "edit_adapter_tiny.pt"))
    vcfg = dict(width=208, layers=2, heads=2, mlp_width=520, output_dim=256, n_queries=256, patch=14)
    rcfg = dict(synth.TINY_RESAMPLER_XL, embedding_dim=256)
    ucfg = dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160, in_channels=8)
    B, hw = 1, 8
    feats = synth.randn("edit_golden_feats", (B, 64, 256))
    noise = synth.randn("edit_golden_noise", (B, 4, hw, hw))
    src = synth.randn("edit_golden_src", (B, 3, hw * 8, hw * 8)).clamp(-1, 1)
    neg = ovit.vit_down(ovit.vit_forward(synth.vit_state_dict(**vcfg), torch.zeros(1, 3, 224, 224), 2))
    prompt, pooled = orx.resampler_xl(synth.resampler_xl_state_dict(rcfg), rcfg, torch.cat([feats, neg.expand(B, -1, -1)]))
    assert rel(prompt[:B], g["prompt"]) < 2e-5 and rel(prompt[B:], g["neg_prompt"]) < 2e-5
    assert rel(pooled[:B], g["pooled"]) < 2e-5 and rel(pooled[B:], g["neg_pooled"]) < 2e-5
    # image-tensor path (adapter_modules.py:100-108): positive and negative are both the un-pooled 256 tokens
    it = synth.image("edit_golden_img224", 1, 224)
    f2 = ovit.vit_forward(synth.vit_state_dict(**vcfg), torch.cat([it, torch.zeros_like(it)]), 2)
    p2, pool2 = orx.resampler_xl(synth.resampler_xl_state_dict(rcfg), rcfg, f2)
    assert rel(p2[:1], g["tensor_prompt"]) < 2e-5 and rel(p2[1:], g["tensor_neg_prompt"]) < 2e-5
    assert rel(pool2[:1], g["tensor_pooled"]) < 2e-5 and rel(pool2[1:], g["tensor_neg_pooled"]) < 2e-5
    assert osd.Euler().set_timesteps(g["steps"]).timesteps.tolist() == g["unet_timesteps"]
    v_sd = synth.vae_state_dict(synth.TINY_VAE)
    il = osd.vae_encode_mode(v_sd, synth.TINY_VAE, src)
    lat = osd.edit_sample(synth.unet_state_dict(ucfg), ucfg, noise, il, prompt[:B], pooled[:B], prompt[B:], pooled[B:], steps=g["steps"], size=hw * 8)
    assert rel(lat, g["latents"]) < 2e-5
    img = osd.vae_decode(v_sd, synth.TINY_VAE, lat / synth.TINY_VAE["scaling_factor"])
    assert rel(img, g["image"]) < 2e-5


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

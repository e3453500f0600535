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

OUT = os.environ.get("SEEDX_GOLDEN_OUT", HERE)          # tests regenerate into a scratch directory and compare with the committed files

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
    torch.save(out, os.path.join(OUT, "vit_small.pt"))
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
    torch.save(out, os.path.join(OUT, "resamplers.pt"))
    print("resamplers", {k: tuple(v.shape) for k, v in out.items()})


def golden_llama():
    """Reference LlamaForCausalLM (xformers stub = exact SDPA) on the tiny config: prefill logits/hidden, 3 cached decode steps,
    and a 40-token greedy loop driven around the reference forward with the reference's own AutoImageTokenGenerationProcessor."""
    from transformers import LlamaConfig
    mod = ref_module("src.models.mllm.modeling_llama_xformer")
    gen = ref_module("src.models.mllm.generation")
    cfg = synth.TINY_LLAMA
    hc = LlamaConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], intermediate_size=cfg["ffn"], num_hidden_layers=cfg["layers"],
                     num_attention_heads=cfg["heads"], rms_norm_eps=cfg["eps"], max_position_embeddings=2048)
    hc.pad_token_id = 0
    model = mod.LlamaForCausalLM(hc).eval()
    sd = synth.llama_state_dict(cfg)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in m for m in missing), (missing, unexpected)
    tok = synth.SynthTokenizer(vocab=cfg["vocab"])
    proc = gen.AutoImageTokenGenerationProcessor(tok, num_img_gen_tokens=64)
    P = 45
    ids = [tok.bos_token_id] + [int(v) for v in (synth.randn("llama_ids", (P - 1,)).abs() * 1000).long() % (tok.base - 3) + 3]
    emb = model.get_input_embeddings()(torch.tensor([ids])).detach()
    emb[0, 5:15] = synth.randn("llama_img_rows", (10, cfg["hidden"]))          # some rows replaced by "image" embeddings
    out = {"ids": ids, "embeds": emb[0].clone()}

    def fwd(**kw):
        with torch.no_grad():
            return model(use_cache=True, output_hidden_states=True, return_dict=True, **kw)

    o = fwd(inputs_embeds=emb, attention_mask=torch.ones(1, P, dtype=torch.long), position_ids=torch.arange(P)[None])
    out["prefill_logits"] = o.logits[0].float()
    out["prefill_hidden"] = o.hidden_states[-1][0].float()
    # greedy loop (HF 4.30 greedy_search semantics) around the reference forward + the reference's processor
    def greedy(ids_, emb_, n_new):
        o_ = fwd(inputs_embeds=emb_, attention_mask=torch.ones(1, len(ids_), dtype=torch.long), position_ids=torch.arange(len(ids_))[None])
        seq, past, logits = list(ids_), o_.past_key_values, o_.logits[:, -1, :]
        gen_ids, hiddens, step_logits = [], [], []
        for step in range(n_new):
            nxt = int(proc(torch.tensor([seq]), logits.clone()).argmax(-1))
            gen_ids.append(nxt)
            seq.append(nxt)
            if step == n_new - 1:
                break
            o_ = fwd(input_ids=torch.tensor([[nxt]]), past_key_values=past, attention_mask=torch.ones(1, len(seq), dtype=torch.long),
                     position_ids=torch.tensor([[len(seq) - 1]]))
            past, logits = o_.past_key_values, o_.logits[:, -1, :]
            hiddens.append(o_.hidden_states[-1][0, -1].float())
            step_logits.append(o_.logits[0, -1].float())
        return gen_ids, torch.stack(hiddens), torch.stack(step_logits)

    g, h, sl = greedy(ids, emb, 16)
    out["text_gen_ids"], out["text_hidden"], out["text_step_logits"] = g, h, sl
    # prompt ending in <img>: the processor forces <img_00000..00063></img>, then free text
    ids_b = ids + [tok.encode("<img>")[0]]
    emb_b = torch.cat([emb, model.get_input_embeddings()(torch.tensor([[ids_b[-1]]])).detach()], dim=1)
    g, h, sl = greedy(ids_b, emb_b, 72)
    out["img_gen_ids"], out["img_hidden"] = g, h
    assert g[:65] == tok.encode("".join("<img_{:05d}>".format(i) for i in range(64)) + "</img>"), g[:66]
    # the reference processor on two hand-made rows (pins the restated processor)
    s = synth.randn("proc_scores", (1, cfg["vocab"]))
    out["proc_in"] = s.clone()
    out["proc_out_text"] = proc(torch.tensor([[5, 6, 7]]), s.clone())
    out["proc_out_img"] = proc(torch.tensor([[5, tok.encode("<img_00010>")[0]]]), s.clone())
    torch.save(out, os.path.join(OUT, "llama_tiny.pt"))
    print("llama_tiny: text", out["text_gen_ids"][:8], "img-span tail", out["img_gen_ids"][62:])


def golden_llama_lora():
    """Reference get_peft_model_with_resize_embedding (src/models/mllm/peft_models.py:27-106) over the reference LlamaForCausalLM with the
    PEFT 0.4.0 vendored under /root/reference/proj/peft: vocabulary grown 1024 -> 1034 (mean-initialised rows), LoRA r=4 alpha=8 on
    the seven projections, modules_to_save norms; UN-MERGED forward in fp32 = golden logits / hidden states."""
    import contextlib
    import types
    import transformers  # noqa: F401  (must be imported before the accelerate stub)
    from transformers import LlamaConfig
    from _ref_import import install_stubs
    install_stubs()

    class _Any(types.ModuleType):
        def __getattr__(self, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return lambda *a, **k: None
    for name in ["accelerate", "accelerate.hooks", "accelerate.utils", "accelerate.utils.modeling", "accelerate.big_modeling"]:
        sys.modules.setdefault(name, _Any(name))
    sys.modules["deepspeed"].zero = types.SimpleNamespace(GatheredParameters=lambda *a, **k: contextlib.nullcontext())
    sys.path.insert(0, "/root/reference/proj/peft/src")
    import peft
    from seedx_b200 import compat
    compat.install()                                        # hydra / omegaconf stand-ins for the reference module's imports
    assert peft.__version__ == "0.4.0" and "/root/reference" in peft.__file__
    pm = ref_module("src.models.mllm.peft_models")
    mod = ref_module("src.models.mllm.modeling_llama_xformer")
    cfg = synth.TINY_LLAMA
    hc = LlamaConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], intermediate_size=cfg["ffn"], num_hidden_layers=cfg["layers"],
                     num_attention_heads=cfg["heads"], rms_norm_eps=cfg["eps"], max_position_embeddings=2048)
    hc.pad_token_id = 0
    hc.tie_word_embeddings = False
    model = mod.LlamaForCausalLM(hc).eval()
    missing, unexpected = model.load_state_dict(synth.llama_state_dict(cfg), strict=False)
    assert not unexpected and all("rotary" in m for m in missing), (missing, unexpected)
    new_vocab = cfg["vocab"] + 10
    lcfg = peft.LoraConfig(r=4, lora_alpha=8, lora_dropout=0.05, task_type="CAUSAL_LM",
                           target_modules=["q_proj", "v_proj", "k_proj", "o_proj", "gate_proj", "down_proj", "up_proj"],
                           modules_to_save=["input_layernorm", "post_attention_layernorm", "norm"])
    pmodel = pm.get_peft_model_with_resize_embedding(model, peft_config=lcfg, vocab_size=new_vocab, torch_dtype="fp32").eval()
    sd = pmodel.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items() if ".lora_" in k or ".modules_to_save." in k}
    assert len(shapes) == 2 * 7 * cfg["layers"] + 2 * cfg["layers"] + 1, len(shapes)
    pmodel.load_state_dict(synth.lora_fixture(shapes), strict=False)
    out = {"peft_keys": list(sd.keys()), "shapes": shapes, "new_vocab": new_vocab, "r": 4, "lora_alpha": 8,
           "embed_new_rows": sd["base_model.model.model.embed_tokens.weight"][cfg["vocab"]:].clone(),
           "head_new_rows": sd["base_model.model.lm_head.weight"][cfg["vocab"]:].clone()}
    P = 40
    ids = [1] + [int(v) for v in (synth.randn("lora_ids", (P - 4,)).abs() * 1000).long() % (cfg["vocab"] - 3) + 3] + [new_vocab - 1, new_vocab - 7, 17]
    out["ids"] = ids
    with torch.no_grad():
        o = pmodel(input_ids=torch.tensor([ids]), attention_mask=torch.ones(1, P, dtype=torch.long), position_ids=torch.arange(P)[None],
                   output_hidden_states=True, return_dict=True, use_cache=False)
    out["logits"], out["hidden"] = o.logits[0].float(), o.hidden_states[-1][0].float()
    torch.save(out, os.path.join(OUT, "llama_lora_tiny.pt"))
    print("llama_lora_tiny: logits", tuple(out["logits"].shape), "keys e.g.", [k for k in out["peft_keys"] if "layers.0.self_attn.q_proj" in k])


def agent_fixture(tok):
    """the request of the agent golden: 2 images (2 + 1 views), three chat turns, prompt forced to open an image span"""
    img = "".join("<img_{:05d}>".format(i) for i in range(64))
    text = "[INST] " + "<patch>" + img + "</patch>" + "<img>" + img + "</img>" + "<img>" + img + "</img>" + "what changed? [/INST]\nthe sky\n[INST] draw it again [/INST]\n<img>"
    ids = torch.tensor([tok.bos_token_id] + tok.encode(text))
    first = tok.tok2id["<img_00000>"]
    lst = ids.tolist()
    starts = [i for i, t in enumerate(lst) if t in (tok.tok2id["<img>"], tok.tok2id["<patch>"])]
    ends = [i for i, t in enumerate(lst) if t in (tok.tok2id["</img>"], tok.tok2id["</patch>"])]
    mask = torch.zeros_like(ids, dtype=torch.bool)
    for a, b in zip(starts, ends):
        mask[a + 1:b] = True
    assert int(mask.sum()) == 3 * 64 and lst[int(mask.nonzero()[0])] == first
    return ids.unsqueeze(0), mask.unsqueeze(0), synth.randn("agent_golden_img", (3, 256, 320)), torch.tensor([[0.25, 0.5], [0.5, 0.5], [0.5, 0.5]])


def golden_agent():
    """The reference's OWN ContinuousLVLM.generate (src/models/mllm/seed_x.py:130-223: input resampler + patch-position row, mask scatter, logits
    processor, hidden-state harvest, output resampler, text assembly) over the reference LlamaForCausalLM and Resamplers.  Only `llm.generate` is
    substituted: the installed transformers (5.x) cannot drive the reference model, so a greedy loop with HF-4.30 `greedy_search` semantics
     runs the reference forward and the reference's logits processor list, and returns `.sequences` / `.hidden_states` in HF's layout."""
    import types
    from transformers import LlamaConfig
    mod = ref_module("src.models.mllm.modeling_llama_xformer")
    sx = ref_module("src.models.mllm.seed_x")
    qv = ref_module("src.models.tokenizer.qwen_visual")
    cfg, vit_dim = synth.TINY_LLAMA, 320
    hc = LlamaConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], intermediate_size=cfg["ffn"], num_hidden_layers=cfg["layers"],
                     num_attention_heads=cfg["heads"], rms_norm_eps=cfg["eps"], max_position_embeddings=2048)
    hc.pad_token_id = 0
    llm = mod.LlamaForCausalLM(hc).eval()
    missing, unexpected = llm.load_state_dict(synth.llama_state_dict(cfg), strict=False)
    assert not unexpected and all("rotary" in m for m in missing), (missing, unexpected)
    agent = sx.ContinuousLVLM(llm=llm, input_resampler=qv.Resampler(grid_size=8, embed_dim=cfg["hidden"], num_heads=2, kv_dim=vit_dim),
                              output_resampler=qv.Resampler(grid_size=8, embed_dim=vit_dim, num_heads=2, kv_dim=cfg["hidden"]),
                              add_patch_pos=True, vit_down=True, mse=True).eval()
    missing, unexpected = agent.load_state_dict(synth.agent_state_dict(cfg["hidden"], vit_dim), strict=False)
    assert not unexpected and all(m.startswith("llm.") for m in missing), (missing, unexpected)

    def fwd(**kw):
        with torch.no_grad():
            return llm(use_cache=True, output_hidden_states=True, return_dict=True, **kw)

    def hf_generate(input_ids=None, inputs_embeds=None, logits_processor=None, max_new_tokens=20, **kw):
        assert kw.get("output_hidden_states") and kw.get("return_dict_in_generate") and kw.get("do_sample") is False and kw.get("num_beams") == 1
        P = input_ids.shape[1]
        o = fwd(inputs_embeds=inputs_embeds, attention_mask=torch.ones(1, P, dtype=torch.long), position_ids=torch.arange(P)[None])
        seq, past, logits = input_ids[0].tolist(), o.past_key_values, o.logits[:, -1, :]
        hidden = [(o.hidden_states[-1],)]
        for step in range(max_new_tokens):
            nxt = int(logits_processor(torch.tensor([seq]), logits.clone().float()).argmax(-1))
            seq.append(nxt)
            if nxt == 2 or step == max_new_tokens - 1:              # eos_token_id of the LLaMA generation config
                break
            o = fwd(input_ids=torch.tensor([[nxt]]), past_key_values=past, attention_mask=torch.ones(1, len(seq), dtype=torch.long),
                    position_ids=torch.tensor([[len(seq) - 1]]))
            past, logits = o.past_key_values, o.logits[:, -1, :]
            hidden.append((o.hidden_states[-1],))
        return types.SimpleNamespace(sequences=torch.tensor([seq]), hidden_states=tuple(hidden))

    llm.generate = hf_generate
    tok = synth.SynthTokenizer(vocab=cfg["vocab"])
    input_ids, ids_cmp_mask, image_embeds, patch_pos = agent_fixture(tok)
    with torch.no_grad():
        res = agent.generate(tokenizer=tok, input_ids=input_ids, image_embeds=image_embeds, embeds_cmp_mask=torch.ones((3, 64), dtype=torch.bool),
                             ids_cmp_mask=ids_cmp_mask, patch_positions=patch_pos, max_new_tokens=70, num_img_gen_tokens=64, device="cpu", dtype=torch.float32)
    assert res["has_img_output"] and res["num_gen_imgs"] == 1
    out = dict(text=res["text"], has_img_output=res["has_img_output"], num_gen_imgs=res["num_gen_imgs"], img_gen_feat=res["img_gen_feat"].float(),
               input_ids=input_ids, ids_cmp_mask=ids_cmp_mask, patch_pos=patch_pos)
    torch.save(out, os.path.join(OUT, "agent_tiny.pt"))
    print("agent_tiny: text", repr(res["text"]), "feat", tuple(out["img_gen_feat"].shape))


def golden_edit_adapter():
    """The reference's OWN SDXLAdapterWithLatentImage.generate (src/models/detokenizer/adapter_modules.py:172-287) driving the reference's OWN
    StableDiffusionXLText2ImageAndEditPipeline.__call__ (pipeline_stable_diffusion_xl_t2i_edit.py:618-994) with the reference's OWN ViT and
    ResamplerXLV2 — negative conditioning, [text, image, uncond] batch, time ids, VAE-encoded source latents, sigma-space 3-way CFG, Euler steps.
    diffusers itself is absent: import-time names come from tests/golden/_ref_import.install_diffusers_stub (no arithmetic), and the three objects
    diffusers would provide (UNet, VAE, scheduler) are thin adapters over oracle/sdxl.py — so this pins everything in the de-tokenizer EXCEPT the
    diffusers-internal UNet / VAE / Euler arithmetic."""
    import types
    from _ref_import import install_diffusers_stub
    from oracle import sdxl as osd
    install_diffusers_stub()
    am = ref_module("src.models.detokenizer.adapter_modules")
    rs = ref_module("src.models.detokenizer.resampler")
    qv = ref_module("src.models.tokenizer.qwen_visual")
    vcfg = dict(VIT_SMALL)
    rcfg = dict(synth.TINY_RESAMPLER_XL, embedding_dim=256)
    ucfg = dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160, in_channels=8)
    u_sd, v_sd = synth.unet_state_dict(ucfg), synth.vae_state_dict(synth.TINY_VAE)
    vit = qv.VisionTransformerWithAttnPool(image_size=448, patch_size=14, width=vcfg["width"], layers=vcfg["layers"], heads=vcfg["heads"],
                                           mlp_ratio=vcfg["mlp_width"] / vcfg["width"], n_queries=256, output_dim=vcfg["output_dim"]).eval()
    vit.load_state_dict({k: v for k, v in synth.vit_state_dict(**vcfg).items() if k != "attn_pool.pos_embed"}, strict=False)
    rx = rs.ResamplerXLV2(normalize=False, **rcfg).eval()
    rx.load_state_dict(synth.resampler_xl_state_dict(rcfg), strict=True)

    class Sched:                                            # diffusers EulerDiscreteScheduler surface over oracle.Euler
        order = 1
        config = types.SimpleNamespace(num_train_timesteps=1000)

        def __init__(self):
            self.e = osd.Euler()

        def set_timesteps(self, n, device=None):
            self.e.set_timesteps(n)
            self.timesteps, self.sigmas, self.init_noise_sigma = self.e.timesteps, self.e.sigmas, self.e.init_noise_sigma

        def _i(self, t):
            return int((self.timesteps == t).nonzero()[0])

        def scale_model_input(self, sample, t):
            return self.e.scale_model_input(sample, self._i(t))

        def step(self, model_output, t, sample, return_dict=False):
            return (self.e.step(model_output, self._i(t), sample),)

    class UNet:                                             # diffusers UNet2DConditionModel surface over oracle.unet_forward
        dtype = torch.float32
        config = types.SimpleNamespace(sample_size=16, addition_time_embed_dim=ucfg["addition_time_embed_dim"], in_channels=8)
        add_embedding = types.SimpleNamespace(linear_1=types.SimpleNamespace(in_features=u_sd["add_embedding.linear_1.weight"].shape[1]))
        calls = []

        def __call__(self, x, t, encoder_hidden_states=None, cross_attention_kwargs=None, added_cond_kwargs=None, return_dict=False):
            self.calls.append(float(t))
            return (osd.unet_forward(u_sd, ucfg, x, float(t), encoder_hidden_states, added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"]),)

    class VAE:                                              # diffusers AutoencoderKL surface over oracle.vae_*
        dtype = torch.float32
        config = types.SimpleNamespace(block_out_channels=synth.TINY_VAE["block_out_channels"], latent_channels=4,
                                       scaling_factor=synth.TINY_VAE["scaling_factor"], force_upcast=False)

        def encode(self, x):
            mode = osd.vae_encode_mode(v_sd, synth.TINY_VAE, x)
            return types.SimpleNamespace(latent_dist=types.SimpleNamespace(mode=lambda: mode))

        def decode(self, z, return_dict=False):
            return (osd.vae_decode(v_sd, synth.TINY_VAE, z),)

    unet = UNet()
    ad = am.SDXLAdapterWithLatentImage(unet=unet, resampler=rx, full_ft=True, set_trainable_late=True, vit_down=True).eval()
    ad.init_pipe(vae=VAE(), scheduler=Sched(), visual_encoder=vit, image_transform=None, dtype=torch.float32, device="cpu")
    B, hw, steps = 1, 8, 3
    feats = synth.randn("edit_golden_feats", (B, 64, 256))
    noise = synth.randn("edit_golden_noise", (B, 4, hw, hw))
    src = synth.randn("edit_golden_src", (B, 3, hw * 8, hw * 8)).clamp(-1, 1)
    out = {}
    with torch.no_grad():
        out["latents"] = ad.generate(image_embeds=feats, latent_image=src, num_inference_steps=steps, height=hw * 8, width=hw * 8, latents=noise.clone(),
                                     input_image_size=224, guidance_scale=7.5, image_guidance_scale=1.5, output_type="latent").float()
        out["image"] = ad.generate(image_embeds=feats, latent_image=src, num_inference_steps=steps, height=hw * 8, width=hw * 8, latents=noise.clone(),
                                   input_image_size=224, guidance_scale=7.5, image_guidance_scale=1.5, output_type="pt").float()
        p, n, pp, npool = ad.get_image_embeds(image_embeds=feats, return_negative=True, image_size=224)
        # reconstruction path (eval_seed_x_detokenizer.py): an image tensor in -> un-pooled 256-token conditioning, negative = ViT(zeros) un-pooled too
        it = synth.image("edit_golden_img224", 1, 224)
        tp, tn, tpp, tnp = ad.get_image_embeds(image_tensor=it, return_negative=True)
    out.update(prompt=p.float(), neg_prompt=n.float(), pooled=pp.float(), neg_pooled=npool.float(), steps=steps, unet_timesteps=unet.calls[:steps],
               tensor_prompt=tp.float(), tensor_neg_prompt=tn.float(), tensor_pooled=tpp.float(), tensor_neg_pooled=tnp.float())
    torch.save(out, os.path.join(OUT, "edit_adapter_tiny.pt"))
    print("edit_adapter_tiny: latents", tuple(out["latents"].shape), "image", tuple(out["image"].shape), "timesteps", out["unet_timesteps"])


def golden_resampler_xl():
    rs = ref_module("src.models.detokenizer.resampler")
    out = {}
    for name, cfg in (("tiny", synth.TINY_RESAMPLER_XL), ("full", synth.RESAMPLER_XL)):
        m = rs.ResamplerXLV2(normalize=False, **cfg).eval()
        m.load_state_dict(synth.resampler_xl_state_dict(cfg), strict=True)
        for n_tok in (64, 256):
            x = synth.randn(f"rxl_{name}_{n_tok}", (2, n_tok, cfg["embedding_dim"]))
            with torch.no_grad():
                p, pooled = m(x)
            out[f"{name}_{n_tok}_prompt"], out[f"{name}_{n_tok}_pooled"] = p.float().half(), pooled.float()
    torch.save(out, os.path.join(OUT, "resampler_xl.pt"))
    print("resampler_xl", {k: tuple(v.shape) for k, v in out.items()})


def golden_preprocess():
    """reference host pre-processing (transforms.get_transform + any_res.process_anyres_image) on seeded noise images: store
    shapes, patch positions and float64 checksums (the tensors themselves are large)."""
    import numpy as np
    from PIL import Image
    ar, tr = ref_module("src.inference.any_res"), ref_module("src.processer.transforms")
    base = 448
    grids = [[base * int(g.split("x")[0]), base * int(g.split("x")[1])] for g in ["1x1", "1x2", "1x3", "2x1", "3x1", "1x4", "4x1", "2x2"]]
    rng = np.random.RandomState(0)
    out = {"grids": grids, "cases": []}
    for (w, h) in [(448, 448), (1024, 1024), (896, 896), (1300, 500), (300, 1200), (640, 480), (2000, 900), (333, 777)]:
        arr = rng.randint(0, 255, (h, w, 3), dtype=np.uint8)
        img = Image.fromarray(arr)
        t = tr.get_transform("clip", keep_ratio=False, image_size=base)
        views, pos = ar.process_anyres_image(img, t, grids, base)
        tk = tr.get_transform("clip", keep_ratio=True, image_size=base)(img)
        out["cases"].append(dict(size=(w, h), shape=tuple(views.shape), pos=pos.clone(), sum=float(views.double().sum()),
                                 abssum=float(views.double().abs().sum()), first=views[0, :, :4, :4].clone(), last=views[-1, :, -4:, -4:].clone(),
                                 keep_sum=float(tk.double().sum())))
    torch.save(out, os.path.join(OUT, "preprocess.pt"))
    print("preprocess", [c["shape"] for c in out["cases"]])


if __name__ == "__main__":
    which = sys.argv[1:] or ["vit", "resamplers", "llama", "llama_lora", "agent", "resampler_xl", "preprocess", "edit_adapter"]
    if "vit" in which:
        golden_vit()
    if "resamplers" in which:
        golden_resamplers()
    if "llama" in which:
        golden_llama()
    if "llama_lora" in which:
        golden_llama_lora()
    if "agent" in which:
        golden_agent()
    if "resampler_xl" in which:
        golden_resampler_xl()
    if "edit_adapter" in which:
        golden_edit_adapter()
    if "preprocess" in which:
        golden_preprocess()

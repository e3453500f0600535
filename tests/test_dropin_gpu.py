"""Drop-in builder surface on the GPU: the reference's `_target_` YAML factories + `from_pretrained` loaders, fed with tiny checkpoints
written in the reference's on-disk layouts (pretrained/QwenViT/*.pt, HF llm dir, agent/adapter pytorch_model.bin, diffusers unet/vae dirs)."""
import json
import os

import pytest
import torch

from seedx_b200 import compat, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_pretrained(root):
    from safetensors.torch import save_file
    vit_cfg = dict(width=208, layers=2, heads=2, mlp_width=520, output_dim=256, n_queries=256, patch=14)
    os.makedirs(root / "QwenViT")
    torch.save(synth.vit_state_dict(**vit_cfg), root / "QwenViT" / "qwen_vit_G.pt")
    lc = synth.TINY_LLAMA
    os.makedirs(root / "seed_x_i" / "llm")
    json.dump(dict(vocab_size=lc["vocab"], hidden_size=lc["hidden"], num_hidden_layers=lc["layers"], num_attention_heads=lc["heads"],
                   intermediate_size=lc["ffn"], rms_norm_eps=lc["eps"]), open(root / "seed_x_i" / "llm" / "config.json", "w"))
    save_file({k: v.half().contiguous() for k, v in synth.llama_state_dict(lc).items()}, str(root / "seed_x_i" / "llm" / "model.safetensors"))
    os.makedirs(root / "seed_x_i" / "agent")
    torch.save(synth.agent_state_dict(lc["hidden"], 256), root / "seed_x_i" / "agent" / "pytorch_model.bin")
    ucfg = dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160)
    sd_dir = root / "stable-diffusion-xl-base-1.0"
    os.makedirs(sd_dir / "unet"), os.makedirs(sd_dir / "vae"), os.makedirs(sd_dir / "scheduler")
    json.dump(dict(in_channels=4, out_channels=4, block_out_channels=list(ucfg["block_out_channels"]), layers_per_block=2,
                   transformer_layers_per_block=list(ucfg["transformer_layers"]), attention_head_dim=list(ucfg["heads"]),
                   cross_attention_dim=256, addition_time_embed_dim=ucfg["addition_time_embed_dim"],
                   down_block_types=["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"]), open(sd_dir / "unet" / "config.json", "w"))
    save_file({k: v.half().contiguous() for k, v in synth.unet_state_dict(ucfg).items()}, str(sd_dir / "unet" / "diffusion_pytorch_model.safetensors"))
    vcfg = synth.TINY_VAE
    json.dump(dict(block_out_channels=list(vcfg["block_out_channels"]), layers_per_block=vcfg["layers_per_block"], latent_channels=4,
                   scaling_factor=0.13025), open(sd_dir / "vae" / "config.json", "w"))
    torch.save(synth.vae_state_dict(vcfg), sd_dir / "vae" / "diffusion_pytorch_model.bin")
    json.dump(dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1), open(sd_dir / "scheduler" / "scheduler_config.json", "w"))
    rcfg = dict(synth.TINY_RESAMPLER_XL, embedding_dim=256)
    os.makedirs(root / "seed_detokenizer" / "first_stage")
    torch.save({"resampler." + k: v for k, v in synth.resampler_xl_state_dict(rcfg).items()}, root / "seed_detokenizer" / "first_stage" / "pytorch_model.bin")
    return vit_cfg, lc, ucfg, vcfg, rcfg


def sd_dir_of(pre):
    return pre / "stable-diffusion-xl-base-1.0"


def test_yaml_factories_and_loaders(tmp_path):
    compat.install()
    import hydra
    from diffusers import AutoencoderKL, EulerDiscreteScheduler, UNet2DConditionModel
    from omegaconf import OmegaConf
    pre = tmp_path / "pretrained"
    vit_cfg, lc, ucfg, vcfg, rcfg = _write_pretrained(pre)
    load = lambda rel: OmegaConf.load(os.path.join(ROOT, "configs", rel))  # noqa: E731

    vit = hydra.utils.instantiate(load("visual_encoder/qwen_vitg_448.yaml"), width=208, layers=2, heads=2, mlp_ratio=2.5, output_dim=256,
                                  pretrained_model_path=str(pre / "QwenViT" / "qwen_vit_G.pt")).eval().to("cuda", dtype=torch.float16)
    llm = hydra.utils.instantiate(load("clm_models/llm_seed_x_i.yaml"), pretrained_model_name_or_path=str(pre / "seed_x_i" / "llm"),
                                  torch_dtype=torch.float16)
    acfg = load("clm_models/agent_seed_x_i.yaml")
    acfg["input_resampler"].update(embed_dim=lc["hidden"], num_heads=2, kv_dim=256)
    acfg["output_resampler"].update(embed_dim=256, num_heads=2, kv_dim=lc["hidden"])
    agent = hydra.utils.instantiate(acfg, llm=llm, pretrained_model_path=str(pre / "seed_x_i" / "agent" / "pytorch_model.bin")).eval().to("cuda")
    sd = str(pre / "stable-diffusion-xl-base-1.0")
    sched = EulerDiscreteScheduler.from_pretrained(sd, subfolder="scheduler")
    vae = AutoencoderKL.from_pretrained(sd, subfolder="vae").to("cuda", dtype=torch.float16)
    unet = UNet2DConditionModel.from_pretrained(sd, subfolder="unet").to("cuda", dtype=torch.float16)
    adcfg = load("sdxl_adapter/sdxl_qwen_vit_resampler_l4_q64_pretrain_no_normalize.yaml")
    adcfg["resampler"].update(dim=rcfg["dim"], depth=rcfg["depth"], heads=rcfg["heads"], embedding_dim=256, output1_dim=rcfg["output1_dim"],
                              output2_dim=rcfg["output2_dim"])
    adapter = hydra.utils.instantiate(adcfg, unet=unet, pretrained_model_path=str(pre / "seed_detokenizer" / "first_stage" / "pytorch_model.bin")).to("cuda").eval()
    transform = hydra.utils.instantiate(load("processer/qwen_448_transform.yaml"))
    discrete = hydra.utils.instantiate(load("discrete_model/discrete_identity.yaml")).to("cuda").eval()
    adapter.init_pipe(vae=vae, scheduler=sched, visual_encoder=vit, image_transform=transform, discrete_model=discrete, dtype=torch.float16, device="cuda")

    # ---- numeric check of every loader: what was read from disk computes what the oracle computes from the same checkpoint tensors
    from oracle import llm as ollm
    from oracle import resampler_xl as orx
    from oracle import sdxl as osd
    from oracle import vit as ovit
    rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()  # noqa: E731
    x_img = synth.image("dropin_img", 1, 224)
    e_vit = rel(vit(x_img.cuda()), ovit.vit_forward(torch.load(pre / "QwenViT" / "qwen_vit_G.pt"), x_img, 2))
    from safetensors.torch import load_file
    lsd = {k: v.float() for k, v in load_file(str(pre / "seed_x_i" / "llm" / "model.safetensors")).items()}
    ids = [1, 17, 254, 99, 512, 33, 8, 640]
    emb = lsd["model.embed_tokens.weight"][torch.tensor(ids)]
    ref_logits, _, _ = ollm.llama_forward(lsd, lc, emb, 0, None)
    logits, _ = llm.logits_all(llm.prefill(emb.cuda()))
    e_llm = rel(logits, ref_logits)
    usd = {k: v.float() for k, v in load_file(str(sd_dir_of(pre) / "unet" / "diffusion_pytorch_model.safetensors")).items()}
    xs, ctx, te = synth.randn("dropin_x", (1, 4, 16, 16)), synth.randn("dropin_ctx", (1, 16, 256)), synth.randn("dropin_te", (1, 160))
    tid = torch.tensor([[256.0, 256.0, 0.0, 0.0, 256.0, 256.0]])
    e_unet = rel(unet(xs.cuda(), 301.0, ctx.cuda(), added_cond_kwargs=dict(text_embeds=te.cuda(), time_ids=tid.cuda())),
                 osd.unet_forward(usd, ucfg, xs, 301.0, ctx, te, tid))
    vsd = torch.load(sd_dir_of(pre) / "vae" / "diffusion_pytorch_model.bin")
    z = synth.randn("dropin_z", (1, 4, 16, 16))
    e_vae = rel(vae.decode(z.cuda()), osd.vae_decode(vsd, vcfg, z))
    rsd = {k[len("resampler."):]: v for k, v in torch.load(pre / "seed_detokenizer" / "first_stage" / "pytorch_model.bin").items()}
    f_in = synth.randn("dropin_feats", (1, 64, 256))
    p_ref, pool_ref = orx.resampler_xl(rsd, rcfg, f_in)
    p_got, pool_got = adapter.resampler(f_in.cuda().half())
    e_rx = max(rel(p_got, p_ref), rel(pool_got, pool_ref))
    print(f"loaders vs oracle on the checkpoint tensors: ViT {e_vit:.2e}, LLaMA logits {e_llm:.2e}, UNet eps {e_unet:.2e}, VAE {e_vae:.2e}, ResamplerXLV2 {e_rx:.2e}")
    assert e_vit < 1e-3 and e_llm < 1e-3 and e_unet < 5e-3 and e_vae < 5e-3 and e_rx < 2e-3, (e_vit, e_llm, e_unet, e_vae, e_rx)
    assert vae.cfg["force_upcast"] is True and vae.stream_scale == 2.0 ** -7        # config.json without the key: diffusers' default (upcast)

    # the flow of eval_img2edit / eval_text2img on a synthetic image
    import numpy as np
    from PIL import Image
    from seedx_b200 import demo
    from seedx_b200.preprocess import process_anyres_image
    tok = synth.SynthTokenizer(vocab=lc["vocab"])
    img = Image.fromarray(np.random.RandomState(0).randint(0, 255, (500, 700, 3), dtype=np.uint8))
    views, patch_pos = process_anyres_image(img, transform, demo.grid_pinpoints(["1x1"]), 448)
    feats = vit(views.to("cuda"))
    assert feats.shape == (2, 256, 256) and feats.dtype == torch.float16
    input_ids, ids_cmp_mask = demo.image_prompt(tok, views.shape[0], "edit it", force_image=True)
    out = agent.generate(tokenizer=tok, input_ids=input_ids, image_embeds=feats, embeds_cmp_mask=torch.ones(2, dtype=torch.bool),
                         patch_positions=patch_pos, ids_cmp_mask=ids_cmp_mask, max_new_tokens=70, num_img_gen_tokens=64)
    assert out["has_img_output"] and out["img_gen_feat"].shape == (1, 64, 256)
    images = adapter.generate(image_embeds=out["img_gen_feat"], num_inference_steps=3, height=256, width=256, seed=7, input_image_size=224)
    assert len(images) == 1 and images[0].size == (256, 256)
    # reconstruction path from a PIL image (eval_seed_x_detokenizer.py): un-pooled 256-token conditioning
    images2 = adapter.generate(image_pil=img, num_inference_steps=2, height=256, width=256, seed=7)
    assert images2[0].size == (256, 256)
    # same seed -> same image (deterministic sampler)
    again = adapter.generate(image_embeds=out["img_gen_feat"], num_inference_steps=3, height=256, width=256, seed=7, input_image_size=224)
    diff = np.abs(np.asarray(images[0]).astype(np.int32) - np.asarray(again[0]).astype(np.int32))
    assert diff.max() == 0, "same seed gave a different image: max |diff| = %d, %d pixels differ" % (diff.max(), (diff > 0).sum())

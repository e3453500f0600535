"""Shared loader for the src/inference/eval_*.py entry points: the reference scripts' common prologue
(src/inference/eval_img2text_seed_x_i.py:66-118) — OmegaConf.load each YAML, hydra.utils.instantiate its `_target_`, wire the pipe."""
import os

import torch

from . import compat

compat.install()
import hydra  # noqa: E402
from omegaconf import OmegaConf  # noqa: E402

BOI_TOKEN, EOI_TOKEN, IMG_TOKEN = "<img>", "</img>", "<img_{:05d}>"
BOP_TOKEN, EOP_TOKEN = "<patch>", "</patch>"
INSTRUCTION = "[INST] {instruction} [/INST]\n"
RESOLUTION_GRIDS = ["1x1", "1x2", "1x3", "2x1", "3x1", "1x4", "4x1", "2x2"]
BASE_RES = 448
DIFFUSION_PATH = "pretrained/stable-diffusion-xl-base-1.0"


def grid_pinpoints(grids=RESOLUTION_GRIDS, base=BASE_RES):
    return [[int(a) * base, int(b) * base] for a, b in (g.split("x") for g in grids)]


def load(variant="seed_x_i", adapter="sdxl_qwen_vit_resampler_l4_q64_pretrain_no_normalize", with_llm=True, edit=False, dtype=torch.float16,
         device="cuda"):
    """returns a dict with tokenizer, image_transform, visual_encoder, agent_model, adapter (any of which may be None)."""
    from diffusers import AutoencoderKL, EulerDiscreteScheduler, UNet2DConditionModel
    inst = lambda p, **kw: hydra.utils.instantiate(OmegaConf.load(p), **kw)  # noqa: E731
    out = {}
    out["image_transform"] = inst("configs/processer/qwen_448_transform.yaml")
    out["visual_encoder"] = inst("configs/visual_encoder/qwen_vitg_448.yaml").eval().to(device, dtype=dtype)
    if with_llm:
        out["tokenizer"] = inst("configs/tokenizer/clm_llama_tokenizer_224loc_anyres.yaml")
        llm = inst(f"configs/clm_models/llm_{variant}.yaml", torch_dtype=dtype)
        out["agent_model"] = inst(f"configs/clm_models/agent_{variant}.yaml", llm=llm).eval().to(device, dtype=dtype)
    noise_scheduler = EulerDiscreteScheduler.from_pretrained(DIFFUSION_PATH, subfolder="scheduler")
    vae = AutoencoderKL.from_pretrained(DIFFUSION_PATH, subfolder="vae").to(device, dtype=dtype)
    unet = UNet2DConditionModel.from_pretrained(DIFFUSION_PATH, subfolder="unet").to(device, dtype=dtype)
    ad = inst(f"configs/sdxl_adapter/{adapter}.yaml", unet=unet).to(device, dtype=dtype).eval()
    if edit:
        ad.init_pipe(vae=vae, scheduler=noise_scheduler, visual_encoder=out["visual_encoder"], image_transform=out["image_transform"], dtype=dtype,
                     device=device)
    else:
        discrete = inst("configs/discrete_model/discrete_identity.yaml").to(device).eval()
        ad.init_pipe(vae=vae, scheduler=noise_scheduler, visual_encoder=out["visual_encoder"], image_transform=out["image_transform"],
                     discrete_model=discrete, dtype=dtype, device=device)
    out["adapter"] = ad
    return out


def image_prompt(tokenizer, n_views, question, n_tokens=64, force_image=False):
    """token layout of SURVEY.md A.2 / eval_img2text_seed_x_i.py:142-162 -> (input_ids [1,P], ids_cmp_mask [1,P])."""
    img = "".join(IMG_TOKEN.format(i) for i in range(n_tokens))
    image_tokens = (BOP_TOKEN + img + EOP_TOKEN) * (n_views - 1) + BOI_TOKEN + img + EOI_TOKEN if n_views else ""
    prompt = INSTRUCTION.format_map({"instruction": image_tokens + question}) + (BOI_TOKEN if force_image else "")
    ids = torch.tensor([tokenizer.bos_token_id] + tokenizer.encode(prompt, add_special_tokens=False))
    starts = {tokenizer.encode(t, add_special_tokens=False)[0] for t in (BOI_TOKEN, BOP_TOKEN)}
    ends = {tokenizer.encode(t, add_special_tokens=False)[0] for t in (EOI_TOKEN, EOP_TOKEN)}
    mask = torch.zeros_like(ids, dtype=torch.bool)
    s_idx = [i for i, t in enumerate(ids.tolist()) if t in starts]
    e_idx = [i for i, t in enumerate(ids.tolist()) if t in ends]
    for a, b in zip(s_idx, e_idx):
        mask[a + 1:b] = True
    return ids.unsqueeze(0), mask.unsqueeze(0)


def save(images, path):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    images[0].save(path)
    print("saved", path)

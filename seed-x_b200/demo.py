"""Shared loader for the src/inference/eval_*.py entry points: the reference scripts' common prologue
(src/inference/eval_img2text_seed_x_i.py:66-118) — OmegaConf.load each YAML, hydra.utils.instantiate its `_target_`, wire the pipe."""
import os
import re

import torch

from . import compat

compat.install()
import hydra  # noqa: E402
from omegaconf import OmegaConf  # noqa: E402

BOI_TOKEN, EOI_TOKEN, IMG_TOKEN = "<img>", "</img>", "<img_{:05d}>"
BOP_TOKEN, EOP_TOKEN = "<patch>", "</patch>"
INSTRUCTION = "[INST] {instruction} [/INST]\n"          # instruction-tuned checkpoints (eval_*_seed_x_i.py:23, eval_img2edit_seed_x_edit.py:27)
# pre-trained (base) checkpoint templates: eval_text2img_seed_x.py:23, eval_img2text_seed_x.py:55-56
BASE_GEN_PROMPT = "{instruction}" + BOI_TOKEN
BASE_QUESTION_PROMPT = "Question: {instruction}\nAnswer:"
BASE_BBOX_PROMPT = "{instruction} [[ <box_start>"
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


def image_prompt(tokenizer, n_views, question, n_tokens=64, force_image=False, template=INSTRUCTION):
    """token layout of SURVEY.md A.2 / eval_img2text_seed_x_i.py:142-162 -> (input_ids [1,P], ids_cmp_mask [1,P]).  `template` holds one
    ``{instruction}`` slot that receives the image tokens followed by the text: INSTRUCTION for the instruction-tuned checkpoints, one of the
    BASE_* templates for the pre-trained one."""
    img = "".join(IMG_TOKEN.format(i) for i in range(n_tokens))
    image_tokens = (BOP_TOKEN + img + EOP_TOKEN) * (n_views - 1) + BOI_TOKEN + img + EOI_TOKEN if n_views else ""
    prompt = template.format_map({"instruction": image_tokens + question}) + (BOI_TOKEN if force_image else "")
    ids = torch.tensor([tokenizer.bos_token_id] + tokenizer.encode(prompt, add_special_tokens=False))
    mask = span_mask(tokenizer, ids)
    return ids.unsqueeze(0), mask.unsqueeze(0)


def chat_prompt(tokenizer, turns, views_per_image=(), n_tokens=64, system_message="", turn_sep="\n", force_image=False):
    """Multi-turn / multi-image prompt in the layout the model was tuned on (/root/reference/src/data/sft_clm.py:216-262):
    ``<s>{system}\n[INST] {image tokens}{q1} [/INST]\n{a1}\n[INST] {q2} [/INST]\n{a2} …``; the image tokens of every input image —
    ``(<patch>64</patch>) x (views-1) + <img>64</img>`` — precede the first question.  ``turns``: [q1, a1, q2, a2, …, qn] (odd length, the
    last user turn is the one to answer).  Returns (input_ids [1,P], ids_cmp_mask [1,P]); the mask selects 64 x sum(views_per_image) rows in
    the order of ``torch.cat([views of image 0, views of image 1, …])``."""
    if len(turns) % 2 != 1:
        raise ValueError("turns must end with a user turn: [q1, a1, ..., qn]")
    img = "".join(IMG_TOKEN.format(i) for i in range(n_tokens))
    image_tokens = "".join((BOP_TOKEN + img + EOP_TOKEN) * (v - 1) + BOI_TOKEN + img + EOI_TOKEN for v in views_per_image)
    text = ""
    if system_message:
        text += system_message if system_message.endswith("\n") else system_message + "\n"
    for i, content in enumerate(turns):
        if i % 2 == 0:
            text += ("" if i == 0 else turn_sep) + INSTRUCTION.format_map({"instruction": (image_tokens if i == 0 else "") + content})
        else:
            text += content
    if force_image:
        text += BOI_TOKEN
    ids = torch.tensor([tokenizer.bos_token_id] + tokenizer.encode(text, add_special_tokens=False))
    return ids.unsqueeze(0), span_mask(tokenizer, ids).unsqueeze(0)


def span_mask(tokenizer, ids):
    """True on the rows strictly between a ``<img>``/``<patch>`` and its closing tag (eval_img2text_seed_x_i.py:155-159)."""
    starts = {tokenizer.encode(t, add_special_tokens=False)[0] for t in (BOI_TOKEN, BOP_TOKEN)}
    ends = {tokenizer.encode(t, add_special_tokens=False)[0] for t in (EOI_TOKEN, EOP_TOKEN)}
    mask = torch.zeros_like(ids, dtype=torch.bool)
    lst = ids.tolist()
    s_idx = [i for i, t in enumerate(lst) if t in starts]
    e_idx = [i for i, t in enumerate(lst) if t in ends]
    for a, b in zip(s_idx, e_idx):
        mask[a + 1:b] = True
    return mask


def extract_box(output_str):
    """``<box_start><loc-x><loc-y><loc-w><loc-h><box_end>`` spans of a grounded answer -> [[x, y, w, h], …] in 224-bin units, or None
    (eval_img2text_seed_x_i.py:39-46)."""
    boxes = re.findall("<box_start>(.*?)<box_end>", output_str)
    return [[int(n) for n in re.findall(r"<loc-(\d+)>", b)] for b in boxes] if boxes else None


def box_to_pixels(bbox, width, height, bins=224):
    """(x_center, y_center, w, h) in `bins` units -> integer pixel corners (x1, y1, x2, y2) (eval_img2text_seed_x_i.py:21-33)."""
    xc, yc, bw, bh = (v / bins * s for v, s in zip(bbox, (width, height, width, height)))
    return int(xc - bw / 2), int(yc - bh / 2), int(xc + bw / 2), int(yc + bh / 2)


def visualize_bbox(image, bboxes, save_path, color=(0, 255, 0), thickness=2):
    """draw the grounded boxes on a PIL image and save it (eval_img2text_seed_x_i.py:16-36; PIL instead of OpenCV)."""
    from PIL import ImageDraw
    img = image.convert("RGB").copy()
    d = ImageDraw.Draw(img)
    for b in bboxes:
        if len(b) != 4:                    # a malformed span (the reference's 4-way unpack would raise here): skip it
            continue
        x1, y1, x2, y2 = box_to_pixels(b, *img.size)
        d.rectangle((min(x1, x2), min(y1, y2), max(x1, x2), max(y1, y2)), outline=color, width=thickness)
    os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
    img.save(save_path)
    return img


def save(images, path):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    images[0].save(path)
    print("saved", path)

"""Instruction-guided editing demo (reference flow: src/inference/eval_img2edit_seed_x_edit.py:97-152): source image -> ViT -> agent
(forced image span) -> edit adapter with the 1024x1024 source as latent_image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (the .project-root marker)
from seedx_b200 import demo  # installs the hydra/omegaconf/pyrootutils/diffusers stand-ins when those packages are absent
import pyrootutils
pyrootutils.setup_root(__file__, indicator=".project-root", pythonpath=True)
import re
import torch
from PIL import Image
from any_res import process_anyres_image

m = demo.load(variant="seed_x_edit", adapter="sdxl_qwen_vit_resampler_l4_q64_full_with_latent_image_pretrain_no_normalize", edit=True)
tok, agent = m["tokenizer"], m["agent_model"]
image = Image.open("demo_images/car.jpg").convert("RGB")
source = image.resize((1024, 1024))
views, patch_pos = process_anyres_image(image, m["image_transform"], demo.grid_pinpoints(["1x1"]), demo.BASE_RES)
input_ids, ids_cmp_mask = demo.image_prompt(tok, views.shape[0], "Make it under the sunset")   # no forced <img>: reference :27,117
with torch.no_grad():
    image_embeds = m["visual_encoder"](views.to("cuda"))
    out = agent.generate(tokenizer=tok, input_ids=input_ids, image_embeds=image_embeds, embeds_cmp_mask=torch.ones(views.shape[0], dtype=torch.bool),
                         patch_positions=patch_pos, ids_cmp_mask=ids_cmp_mask, max_new_tokens=512, num_img_gen_tokens=64)
    if out["has_img_output"]:
        images = m["adapter"].generate(image_embeds=out["img_gen_feat"], latent_image=source, num_inference_steps=50)
        demo.save(images, "vis/car_edit.jpg")
print(out["text"])

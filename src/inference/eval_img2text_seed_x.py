"""Image comprehension demo (reference flow: src/inference/eval_img2text_seed_x.py:131-179): any-res tiles -> ViT -> agent.generate -> text."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (the .project-root marker)
from seedx_b200 import demo  # installs the hydra/omegaconf/pyrootutils/diffusers stand-ins when those packages are absent
import pyrootutils
pyrootutils.setup_root(__file__, indicator=".project-root", pythonpath=True)
import re
import torch
from PIL import Image
from any_res import process_anyres_image

m = demo.load(variant="seed_x")
tok, agent = m["tokenizer"], m["agent_model"]
image = Image.open("demo_images/cat_dog.jpeg").convert("RGB")
views, patch_pos = process_anyres_image(image, m["image_transform"], demo.grid_pinpoints(), demo.BASE_RES)
input_ids, ids_cmp_mask = demo.image_prompt(tok, views.shape[0], "Describe this image briefly.", template=demo.BASE_QUESTION_PROMPT)  # reference :55,149-150
with torch.no_grad():
    image_embeds = m["visual_encoder"](views.to("cuda"))
    out = agent.generate(tokenizer=tok, input_ids=input_ids, image_embeds=image_embeds, embeds_cmp_mask=torch.ones(views.shape[0], dtype=torch.bool),
                         patch_positions=patch_pos, ids_cmp_mask=ids_cmp_mask, max_new_tokens=512, num_img_gen_tokens=64)
print(re.sub("<[^>]*>", "", out["text"]))

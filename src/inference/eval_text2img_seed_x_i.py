"""Text-to-image demo (reference flow: src/inference/eval_text2img_seed_x_i.py:82-93): prompt -> agent.generate -> img_gen_feat -> adapter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (the .project-root marker)
from seedx_b200 import demo  # installs the hydra/omegaconf/pyrootutils/diffusers stand-ins when those packages are absent
import pyrootutils
pyrootutils.setup_root(__file__, indicator=".project-root", pythonpath=True)
import re
import torch
from PIL import Image
from any_res import process_anyres_image

m = demo.load(variant="seed_x_i")
tok, agent = m["tokenizer"], m["agent_model"]
caption = "A photo of an astronaut riding a horse on the moon."
input_ids, _ = demo.image_prompt(tok, 0, "Generate an image: " + caption)             # the model opens the <img> span itself (reference :23)
with torch.no_grad():
    out = agent.generate(tokenizer=tok, input_ids=input_ids, num_img_gen_tokens=64)
    if out["has_img_output"]:
        images = m["adapter"].generate(image_embeds=out["img_gen_feat"], num_inference_steps=50)
        demo.save(images, "vis/text2img.jpg")
print(out["text"])

"""De-tokenizer reconstruction demo (reference flow: src/inference/eval_seed_x_detokenizer.py:58-61): PIL image -> ViT -> SDXL adapter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (the .project-root marker)
from seedx_b200 import demo  # installs the hydra/omegaconf/pyrootutils/diffusers stand-ins when those packages are absent
import pyrootutils
pyrootutils.setup_root(__file__, indicator=".project-root", pythonpath=True)
import re
import torch
from PIL import Image
from any_res import process_anyres_image

m = demo.load(with_llm=False)
image = Image.open("demo_images/man.jpg").convert("RGB")
with torch.no_grad():
    images = m["adapter"].generate(image_pil=image, num_inference_steps=50)
demo.save(images, "vis/men_recon.jpg")

"""Conditioned de-tokenizer demo (reference flow: src/inference/eval_seed_x_detokenizer_with_condition.py:58-64):
target image -> ViT features, condition image (1024x1024) -> VAE latents concatenated into the 8-channel UNet."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (the .project-root marker)
from seedx_b200 import demo  # installs the hydra/omegaconf/pyrootutils/diffusers stand-ins when those packages are absent
import pyrootutils
pyrootutils.setup_root(__file__, indicator=".project-root", pythonpath=True)
import re
import torch
from PIL import Image
from any_res import process_anyres_image

m = demo.load(adapter="sdxl_qwen_vit_resampler_l4_q64_full_with_latent_image_pretrain_no_normalize", with_llm=False, edit=True)
target = Image.open("demo_images/bank.png").convert("RGB")
condition = Image.open("demo_images/bank.png").convert("RGB").resize((1024, 1024))
with torch.no_grad():
    images = m["adapter"].generate(image_pil=target, latent_image=condition, num_inference_steps=50)
demo.save(images, "vis/bank_recon.png")

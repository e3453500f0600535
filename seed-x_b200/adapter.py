"""De-tokenizer hosts: ``SDXLAdapter`` (t2i / reconstruction) and ``SDXLAdapterWithLatentImage`` (edit / conditioned) with the
public surface of /root/reference/src/models/detokenizer/adapter_modules.py:11-287 — ``from_pretrained(unet=, resampler=, ...)``,
``init_pipe(vae, scheduler, visual_encoder, image_transform[, discrete_model], dtype, device)``, ``get_image_embeds``,
``generate(image_pil|image_tensor|image_embeds[, latent_image], seed, height, width, guidance_scale, num_inference_steps,
input_image_size) -> list[PIL.Image]``.

Differences that are deliberate: the negative conditioning ViT(zeros) is input independent and is
computed once per image size instead of on every call (D.8/§3.3); ``image_embeds`` may carry a batch (D.9: the reference's
``chunk(2)`` is only right for batch 1); ``latents=`` may be passed for reproducible sampling (D.10).
"""
import numpy as np
import torch
from PIL import Image

from . import ops, trace
from ._lib import SeedxError
from .sampler import DenoiseLoop, decode_to_uint8


class SDXLAdapter:
    branches = 2

    def __init__(self, unet, resampler, full_ft=False, vit_down=False, **kw):
        self.unet, self.resampler = unet, resampler
        self.full_ft, self.vit_down = full_ft, vit_down
        self.device = torch.device("cuda")
        self._neg_cache = {}
        self._loops = {}

    @classmethod
    def from_pretrained(cls, unet, resampler, pretrained_model_path=None, **kwargs):
        kwargs.pop("set_trainable_late", None)
        model = cls(unet=unet, resampler=resampler, **kwargs)
        if pretrained_model_path is not None:
            model.load_state_dict(torch.load(pretrained_model_path, map_location="cpu"))
        return model

    def load_state_dict(self, sd, strict=False):
        """checkpoint keys: 'resampler.*' (always) and 'unet.*' (full fine-tune checkpoints, adapter_modules.py:62-65)."""
        rs = {k[len("resampler."):]: v for k, v in sd.items() if k.startswith("resampler.")}
        if rs:
            self.resampler.load_state_dict(rs)
        un = {k[len("unet."):]: v for k, v in sd.items() if k.startswith("unet.")}
        if un:
            self.unet.load_state_dict(un)
        return [], []

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def init_pipe(self, vae, scheduler, visual_encoder, image_transform, discrete_model=None, dtype=torch.float16, device="cuda"):
        self.vae, self.scheduler = vae, scheduler
        self.visual_encoder = visual_encoder
        self.image_transform = image_transform
        self.discrete_model = discrete_model
        self.dtype = dtype

    # ---- conditioning --------------------------------------------------------------------------------------------------
    def _negative_feats(self, image_size, pooled):
        key = (image_size, pooled)
        if key not in self._neg_cache:
            z = torch.zeros((1, 3, image_size, image_size), device=self.device, dtype=torch.float32)
            f = self.visual_encoder(z)                                           # [1,256,4096]
            if pooled:
                f = ops.avgpool_tokens(f.contiguous(), 4)                        # vit_down: mean over groups of 4 tokens (adapter_modules.py:112-115)
            self._neg_cache[key] = f
        return self._neg_cache[key]

    def get_image_embeds(self, image_pil=None, image_tensor=None, image_embeds=None, return_negative=True, image_size=448):
        assert int(image_pil is not None) + int(image_tensor is not None) + int(image_embeds is not None) == 1
        if image_pil is not None:
            image_tensor = self.image_transform(image_pil).unsqueeze(0)
        if image_tensor is not None:
            feats = self.visual_encoder(image_tensor.to(self.device))            # un-pooled 256 tokens on this path (D.8)
            neg = self._negative_feats(image_tensor.shape[-1], False) if return_negative else None
        else:
            feats = image_embeds.to(self.device)
            neg = self._negative_feats(image_size, self.vit_down) if return_negative else None
        if self.discrete_model is not None:
            feats = self.discrete_model.encode_image_embeds(feats)
        B = feats.shape[0]
        if return_negative:
            allf = torch.empty((2 * B,) + tuple(feats.shape[1:]), device=self.device, dtype=torch.float16)
            ops.unary_f16(feats.reshape(B * feats.shape[1], -1).contiguous(), out=allf[:B].view(B * feats.shape[1], -1))
            for b in range(B):
                ops.unary_f16(neg.reshape(neg.shape[1], -1).contiguous(), out=allf[B + b])
            prompt, pooled = self.resampler(allf)
            return prompt[:B], prompt[B:], pooled[:B], pooled[B:]
        prompt, pooled = self.resampler(feats)
        return prompt, None, pooled, None

    # ---- sampling ------------------------------------------------------------------------------------------------------
    def _loop(self, B, h, w):
        key = (B, h, w)
        if key not in self._loops:
            self._loops[key] = DenoiseLoop(self.unet, self.scheduler, B, (h // 8, w // 8), self.branches, use_graph=True)
        return self._loops[key]

    def _noise(self, B, h, w, seed, latents):
        if latents is not None:
            return latents.to(self.device, torch.float32)
        gen = torch.Generator(self.device).manual_seed(seed) if seed is not None else None
        return torch.randn((B, 4, h // 8, w // 8), generator=gen, device=self.device, dtype=torch.float16).float()  # randn_tensor(dtype=fp16)

    @staticmethod
    def _to_pil(u8):
        arr = u8.cpu().numpy()
        return [Image.fromarray(np.ascontiguousarray(a)) for a in arr]

    def generate(self, image_pil=None, image_tensor=None, image_embeds=None, seed=None, height=1024, width=1024, guidance_scale=7.5,
                 num_inference_steps=30, input_image_size=448, latents=None, output_type="pil", **kwargs):
        p, n, pp, npool = self.get_image_embeds(image_pil=image_pil, image_tensor=image_tensor, image_embeds=image_embeds,
                                               return_negative=True, image_size=input_image_size)
        B = p.shape[0]
        trace.mark("detok.resampler_xl")
        loop = self._loop(B, height, width)
        tid = torch.tensor([[height, width, 0, 0, height, width]], dtype=torch.float32, device=self.device).repeat(2 * B, 1)
        loop.set_condition(torch.cat([n, p]), torch.cat([npool, pp]), tid)        # batch order [negative, positive]
        trace.mark("detok.prepare_cond")
        lat = loop.run(self._noise(B, height, width, seed, latents), steps=num_inference_steps, guidance=guidance_scale)
        trace.mark("detok.denoise_loop")
        if output_type == "latent":
            return lat.clone()
        u8 = decode_to_uint8(self.vae, lat)
        trace.mark("detok.vae_decode")
        return u8 if output_type == "uint8" else self._to_pil(u8)


class SDXLAdapterWithLatentImage(SDXLAdapter):
    """edit variant: 8-channel conv_in (4 noisy latents + 4 source-image latents), 3-way CFG in sigma space
    (adapter_modules.py:172-287 -> pipeline_stable_diffusion_xl_t2i_edit.py:618-994)."""
    branches = 3

    def __init__(self, unet, resampler, full_ft=False, set_trainable_late=False, vit_down=False, **kw):
        super().__init__(unet, resampler, full_ft=full_ft, vit_down=vit_down)
        # set_trainable() widens conv_in to 8 input channels at construction: the first 4 keep the SDXL weights, the 4 image-latent channels
        # start at zero (adapter_modules.py:183-198).  The packed conv_in already pads its input channels to 64 with zeros and the sampler's
        # UNet input buffer always carries 8 channels, so the widened layer is the one in memory: only the declared width changes.  A
        # fine-tuned checkpoint ('unet.*' keys) then replaces the whole UNet in load_state_dict.
        if getattr(unet, "_loaded", False) and unet.cfg.get("in_channels") == 4:
            unet.cfg["in_channels"] = 8

    def init_pipe(self, vae, scheduler, visual_encoder, image_transform, dtype=torch.float16, device="cuda"):
        super().init_pipe(vae, scheduler, visual_encoder, image_transform, None, dtype, device)

    def _image_latents(self, latent_image, height, width):
        """VaeImageProcessor.preprocess ([0,1] -> 2x-1, NCHW) + vae.encode(...).latent_dist.mode(), un-scaled (:523)."""
        if isinstance(latent_image, Image.Image):
            w, h = latent_image.size
            w, h = w - w % 8, h - h % 8
            if (w, h) != latent_image.size:
                latent_image = latent_image.resize((w, h), Image.LANCZOS)
            a = np.asarray(latent_image.convert("RGB"), dtype=np.uint8).astype(np.float32) / 255.0
            img = torch.from_numpy(a.transpose(2, 0, 1)[None]) * 2.0 - 1.0
        else:
            img = latent_image
        if img.shape[1] == 4:
            return img.to(self.device, torch.float32)
        return self.vae.encode_mode(img)

    def generate(self, image_pil=None, image_tensor=None, image_embeds=None, latent_image=None, seed=42, height=1024, width=1024,
                 guidance_scale=7.5, num_inference_steps=30, input_image_size=448, image_guidance_scale=1.5, latents=None,
                 output_type="pil", **kwargs):
        if self.unet.cfg["in_channels"] != 8:
            raise SeedxError("SDXLAdapterWithLatentImage needs the 8-channel conv_in checkpoint (adapter_modules.py:183-198)")
        p, n, pp, npool = self.get_image_embeds(image_pil=image_pil, image_tensor=image_tensor, image_embeds=image_embeds,
                                               return_negative=True, image_size=input_image_size)
        B = p.shape[0]
        trace.mark("detok.resampler_xl")
        loop = self._loop(B, height, width)
        tid = torch.tensor([[height, width, 0, 0, height, width]], dtype=torch.float32, device=self.device).repeat(3 * B, 1)
        il = self._image_latents(latent_image, height, width) if latent_image is not None else \
            torch.zeros((B, 4, height // 8, width // 8), device=self.device)
        if il.shape[0] == 1 and B > 1:
            il = il.expand(B, -1, -1, -1).contiguous()
        trace.mark("detok.vae_encode")
        loop.set_condition(torch.cat([p, n, n]), torch.cat([pp, npool, npool]), tid, image_latents=il)   # [text, image, uncond] (:884-886)
        trace.mark("detok.prepare_cond")
        lat = loop.run(self._noise(B, height, width, seed, latents), steps=num_inference_steps, guidance=guidance_scale,
                       image_guidance=image_guidance_scale)
        trace.mark("detok.denoise_loop")
        if output_type == "latent":
            return lat.clone()
        u8 = decode_to_uint8(self.vae, lat)
        trace.mark("detok.vae_decode")
        return u8 if output_type == "uint8" else self._to_pil(u8)


class DiscreteModleIdentity:
    """src/models/tokenizer/discrete_models.py:7-17 — identity 'discrete' model passed to init_pipe."""

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def encode_image_embeds(self, image_embeds):
        return image_embeds

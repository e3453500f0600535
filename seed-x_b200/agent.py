"""Stage 2 glue host: ``ContinuousLVLM`` (the MLLM agent) — same surface as the reference class
``src.models.mllm.seed_x.ContinuousLVLM`` (/root/reference/src/models/mllm/seed_x.py:22-234): ``from_pretrained(llm=, input_resampler=,
output_resampler=, add_patch_pos=, ...)``, ``generate(tokenizer=, prompt=|input_ids=, image_embeds=, embeds_cmp_mask=, ids_cmp_mask=,
patch_positions=, max_new_tokens=, num_img_gen_tokens=) -> {'text','has_img_output','img_gen_feat','num_gen_imgs'}``.
The training ``forward`` (losses, seed_x.py:48-128) is out of scope.
"""
import torch

from . import ops, trace
from ._lib import SeedxError
from .vit import ResamplerWeights

BOI_TOKEN = "<img>"
EOI_TOKEN = "</img>"
IMG_TOKEN = "<img_{:05d}>"


class Resampler:
    """Config-side stand-in for ``src.models.tokenizer.qwen_visual.Resampler`` (built by the YAML `_target_`): holds the
    hyper-parameters; device weights are packed by ContinuousLVLM.load_state_dict."""

    def __init__(self, grid_size, embed_dim, num_heads, kv_dim=None, **kw):
        self.grid_size, self.embed_dim, self.num_heads, self.kv_dim = grid_size, embed_dim, num_heads, kv_dim


class ContinuousLVLM:
    def __init__(self, llm, input_resampler, output_resampler, lm_loss_scale=1.0, rec_loss_scale=1.0, add_patch_pos=False, vit_down=False,
                 mse=False, **kw):
        self.llm = llm
        self.in_cfg, self.out_cfg = input_resampler, output_resampler
        self.add_patch_pos = add_patch_pos
        self.vit_down = vit_down
        self.device = llm.device
        self._loaded = False

    @classmethod
    def from_pretrained(cls, llm, input_resampler, output_resampler, pretrained_model_path=None, **kw):
        m = cls(llm=llm, input_resampler=input_resampler, output_resampler=output_resampler, **kw)
        if pretrained_model_path is not None:
            m.load_state_dict(torch.load(pretrained_model_path, map_location="cpu"))
        return m

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def load_state_dict(self, sd, strict=False):
        dev = self.device
        self.input_resampler = ResamplerWeights(sd, "input_resampler.", self.in_cfg.num_heads, 1e-5, dev)
        self.output_resampler = ResamplerWeights(sd, "output_resampler.", self.out_cfg.num_heads, 1e-5, dev)
        if self.add_patch_pos:
            w = sd["patch_pos_embed"].float()                     # [4, D]; x @ W == gemm(x, W^T); K padded 4 -> 8 for 16-byte rows
            wt = torch.zeros((w.shape[1], 8), dtype=torch.float16, device=w.device)
            wt[:, :4] = w.t().to(torch.float16)
            self.patch_w = wt.to(dev)
        # a fine-tuned agent checkpoint also carries the LLM's trained tensors under 'llm.' — LoRA pairs, modules_to_save norms,
        # embeddings (README.md:150-160; the reference's load_state_dict(strict=False) loads them into the wrapped model, utils.py:25-42)
        llm_sd = {k[len("llm."):]: v for k, v in sd.items() if k.startswith("llm.")}
        unexpected = self.llm.apply_peft_state_dict(llm_sd) if llm_sd else []
        self._loaded = True
        return [], ["llm." + k for k in unexpected]

    def encode_images(self, image_embeds, patch_positions):
        """input_resampler(image_embeds) + patch-position embedding (seed_x.py:164-171): [N,256,4096] -> fp32 [N*64, D]."""
        N, T, C = image_embeds.shape
        x16 = ops.unary_f16(image_embeds.reshape(N * T, C).to(self.device).contiguous())
        rel = None
        if self.add_patch_pos:
            if patch_positions is None:
                raise SeedxError("add_patch_pos=True needs patch_positions")
            pp = patch_positions.to(self.device, torch.float32)
            pin = torch.zeros((N, 8), device=self.device, dtype=torch.float16)
            ops.unary_f16(pp, out=pin[:, 0:2])
            ops.unary_f16((1.0 - pp), out=pin[:, 2:4])
            rel = ops.gemm(pin, self.patch_w, alpha=0.5, out_dtype=torch.float32)       # [N, D] = cat[p, 1-p]/2 @ W
        r = self.input_resampler
        return self._resample(r, x16, N, T, rel)

    @staticmethod
    def _resample(r, x16, batch, n_kv, bias_g=None):
        """Resampler forward with an optional per-view additive row (fused into the out-projection epilogue)."""
        E, H, Nq = r.embed_dim, r.heads, r.n_queries
        d = E // H
        kv = ops.gemm(x16, r.kv_proj, out_dtype=torch.float32) if r.kv_proj is not None else x16
        v_in, k_in = ops.layernorm(kv, r.ln_kv[0], r.ln_kv[1], r.eps, add=r.pos_k(n_kv))
        k = ops.gemm(k_in, r.wk, bias=r.bk)
        v = ops.gemm(v_in, r.wv, bias=r.bv)
        q = r.q_proj()
        o = torch.empty((batch * Nq, E), device=x16.device, dtype=torch.float16)
        ops.attention(q.view(1, Nq, H, d).permute(0, 2, 1, 3), k.view(batch, n_kv, H, d).permute(0, 2, 1, 3),
                      v.view(batch, n_kv, H, d).permute(0, 2, 1, 3), o.view(batch, Nq, H, d).permute(0, 2, 1, 3), scale=d ** -0.5)
        return ops.gemm(o, r.wo, bias=r.bo, out_dtype=torch.float32, bias_g=bias_g, bias_g_rows=Nq if bias_g is not None else 0)

    def _embed_request(self, input_ids, image_embeds, embeds_cmp_mask, ids_cmp_mask, patch_positions):
        """prompt embeddings with the resampled image features scattered into the <img_k> rows (seed_x.py:157-173)"""
        ids = torch.as_tensor(input_ids).reshape(-1).cpu()
        P = ids.numel()
        D = self.llm.cfg["hidden"]
        x = self.llm.get_input_embeddings()(ids).view(P, D)
        if image_embeds is not None:
            assert embeds_cmp_mask is not None and ids_cmp_mask is not None
            lm = self.encode_images(image_embeds, patch_positions)                  # [N*64, D] fp32
            dst = torch.nonzero(torch.as_tensor(ids_cmp_mask).reshape(-1).cpu()).reshape(-1).to(torch.int32).to(self.device)
            em = torch.as_tensor(embeds_cmp_mask).cpu()
            if em.dim() == 1:                      # one flag per view (eval_img2text_seed_x_i.py:134) -> all 64 rows of that view
                em = em[:, None].expand(-1, lm.shape[0] // em.shape[0])
            src = torch.nonzero(em.reshape(-1)).reshape(-1).to(torch.int32).to(self.device)
            if dst.numel() != src.numel():
                raise SeedxError("ids_cmp_mask and embeds_cmp_mask select different numbers of rows")
            ops.scatter_rows(lm, dst, x, src_idx=src)
        return ids, x

    def _harvest(self, tokenizer, out, P, num_img_gen_tokens):
        """split generated ids into text / image spans and run the output resampler on the harvested hidden rows (seed_x.py:191-223)"""
        D = self.llm.cfg["hidden"]
        gen = out.sequences[0][P:]
        boi = tokenizer.encode(BOI_TOKEN, add_special_tokens=False)[0]
        eoi = tokenizer.encode(EOI_TOKEN, add_special_tokens=False)[0]
        eoi_idx = torch.where(gen == eoi)[0].tolist()
        text_mask = torch.ones_like(gen, dtype=torch.bool)
        feat = None
        if eoi_idx:
            hid = out.last_hidden_states
            rows = torch.empty((len(eoi_idx) * num_img_gen_tokens, D), device=self.device, dtype=torch.float16)
            for j, e in enumerate(eoi_idx):
                if e - num_img_gen_tokens < 0 or e > hid.shape[0]:
                    raise SeedxError("image span is truncated by max_new_tokens")
                ops.unary_f16(hid[e - num_img_gen_tokens:e], out=rows[j * num_img_gen_tokens:(j + 1) * num_img_gen_tokens])
                text_mask[e - num_img_gen_tokens:e] = False
            r = self.output_resampler
            feat = self._resample(r, rows, len(eoi_idx), num_img_gen_tokens).view(len(eoi_idx), r.n_queries, r.embed_dim)
        text_mask[gen == boi] = False
        text = tokenizer.decode(gen[text_mask], skip_special_tokens=False)
        return {"text": text, "has_img_output": bool(eoi_idx), "img_gen_feat": feat, "num_gen_imgs": len(eoi_idx), "ids": gen.tolist()}

    def generate_batch(self, tokenizer, requests, num_img_gen_tokens=64, max_new_tokens=120, suppress_eos=False):
        """Several independent requests decoded in lock-step (the LLM weights are read once per step for all of them).
        requests: list of dicts with the per-request keyword arguments of generate()."""
        if not self._loaded:
            raise SeedxError("ContinuousLVLM: weights not loaded")
        img_str = "".join([BOI_TOKEN] + [IMG_TOKEN.format(i) for i in range(num_img_gen_tokens)] + [EOI_TOKEN])
        img_ids = tokenizer.encode(img_str, add_special_tokens=False)
        eos = getattr(tokenizer, "eos_token_id", None)
        results = []
        for c0 in range(0, len(requests), 8):
            chunk = requests[c0:c0 + 8]
            pairs = [self._embed_request(r.get("input_ids"), r.get("image_embeds"), r.get("embeds_cmp_mask"), r.get("ids_cmp_mask"),
                                         r.get("patch_positions")) for r in chunk]
            trace.mark("llm.embed+input_resampler")
            outs = self.llm.generate_greedy_batch([p[0] for p in pairs], [p[1] for p in pairs], img_ids=img_ids, max_new_tokens=max_new_tokens,
                                                  eos_id=eos, suppress_eos=suppress_eos)
            results += [self._harvest(tokenizer, o, p[0].numel(), num_img_gen_tokens) for o, p in zip(outs, pairs)]
            trace.mark("llm.harvest+output_resampler")
        return results

    def generate(self, tokenizer, prompt=None, input_ids=None, image_embeds=None, embeds_cmp_mask=None, ids_cmp_mask=None,
                 logits_processor=None, num_img_gen_tokens=64, temperature=0.7, num_beams=1, max_new_tokens=120, top_p=0.5,
                 dtype=torch.float16, device="cuda", patch_positions=None, suppress_eos=False):
        """reference signature (seed_x.py:130-145); temperature / top_p / num_beams are accepted and ignored exactly as the reference
        ignores them (do_sample=False greedy search, Appendix D.13)."""
        if prompt is not None:
            input_ids = tokenizer(prompt, return_tensors="pt").input_ids
        req = dict(input_ids=input_ids, image_embeds=image_embeds, embeds_cmp_mask=embeds_cmp_mask, ids_cmp_mask=ids_cmp_mask,
                   patch_positions=patch_positions)
        return self.generate_batch(tokenizer, [req], num_img_gen_tokens=num_img_gen_tokens, max_new_tokens=max_new_tokens,
                                   suppress_eos=suppress_eos)[0]

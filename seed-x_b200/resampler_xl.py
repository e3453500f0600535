"""De-tokenizer front host: ``ResamplerXLV2`` perceiver resampler through libseedx.so — same constructor keywords and
state-dict keys as ``src.models.detokenizer.resampler.ResamplerXLV2`` (/root/reference/src/models/detokenizer/resampler.py:226-286)."""
import torch

from . import ops
from ._lib import SeedxError


class ResamplerXLV2:
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output1_dim=768, output2_dim=1280,
                 ff_mult=4, normalize=True):
        if normalize:
            raise SeedxError("normalize=True is not used by any shipped config (sdxl_adapter/*.yaml) and is not implemented")
        if dim_head not in (64, 128):
            raise SeedxError("dim_head must be 64 or 128")
        self.cfg = dict(dim=dim, depth=depth, dim_head=dim_head, heads=heads, num_queries=num_queries, embedding_dim=embedding_dim,
                        output1_dim=output1_dim, output2_dim=output2_dim, ff_mult=ff_mult)
        self.device = torch.device("cuda")
        self._loaded = False

    def load_state_dict(self, sd, strict=False, prefix=""):
        dev, c = self.device, self.cfg
        h = lambda k: sd[prefix + k].to(dev, torch.float16).contiguous()  # noqa: E731
        f = lambda k: sd[prefix + k].float().to(dev).contiguous()  # noqa: E731
        self.latents = f("latents").view(c["num_queries"], c["dim"])
        self.proj_in = (h("proj_in.weight"), f("proj_in.bias"))
        self.norm_out = (f("norm_out.weight"), f("norm_out.bias"))
        self.layers = []
        for i in range(c["depth"]):
            a, ff = f"layers.{i}.0", f"layers.{i}.1"
            self.layers.append(dict(n1=(f(a + ".norm1.weight"), f(a + ".norm1.bias")), n2=(f(a + ".norm2.weight"), f(a + ".norm2.bias")),
                                    wq=h(a + ".to_q.weight"), wkv=h(a + ".to_kv.weight"), wo=h(a + ".to_out.weight"),
                                    nf=(f(ff + ".0.weight"), f(ff + ".0.bias")), w1=h(ff + ".1.weight"), w2=h(ff + ".3.weight")))
        self.w_out = torch.cat([sd[prefix + "unet_proj_1.weight"], sd[prefix + "unet_proj_2.weight"]], 0).to(dev, torch.float16).contiguous()
        self.b_out = torch.cat([sd[prefix + "unet_proj_1.bias"], sd[prefix + "unet_proj_2.bias"]], 0).float().to(dev).contiguous()
        p = "unet_attnpool."
        self.pool_pos = f(p + "positional_embedding")
        self.pool = dict(wq=h(p + "q_proj.weight"), bq=f(p + "q_proj.bias"), wk=h(p + "k_proj.weight"), bk=f(p + "k_proj.bias"),
                         wv=h(p + "v_proj.weight"), bv=f(p + "v_proj.bias"), wc=h(p + "c_proj.weight"), bc=f(p + "c_proj.bias"))
        self._loaded = True
        return [], []

    def __call__(self, x):
        """x: [B, n, embedding_dim] (device, fp16/fp32) -> (prompt_embeds fp32 [B, Q, out1+out2], pooled fp32 [B, out2])."""
        if not self._loaded:
            raise SeedxError("ResamplerXLV2: weights not loaded")
        c = self.cfg
        B, n, E = x.shape
        D, H, d, Q = c["dim"], c["heads"], c["dim_head"], c["num_queries"]
        inner = H * d
        dev = self.device
        x16 = ops.unary_f16(x.reshape(B * n, E).to(dev).contiguous())
        xf = ops.gemm(x16, self.proj_in[0], bias=self.proj_in[1], out_dtype=torch.float32)              # [B*n, D]
        lat = self.latents.unsqueeze(0).repeat(B, 1, 1).view(B * Q, D)      # a COPY of the learned latents (updated in place below), also for B=1
        kvin = torch.empty((B, n + Q, D), device=dev, dtype=torch.float16)
        lbuf = torch.empty((B * Q, D), device=dev, dtype=torch.float16)
        o = torch.empty((B * Q, inner), device=dev, dtype=torch.float16)
        for L in self.layers:
            ops.layernorm(lat, L["n2"][0], L["n2"][1], 1e-5, out=lbuf)
            for b in range(B):                                                                          # kv input = cat(LN1(x), LN2(latents))
                ops.layernorm(xf[b * n:(b + 1) * n], L["n1"][0], L["n1"][1], 1e-5, out=kvin[b, :n])
                ops.unary_f16(lbuf[b * Q:(b + 1) * Q], out=kvin[b, n:])
            q = ops.gemm(lbuf, L["wq"])
            kv = ops.gemm(kvin.view(B * (n + Q), D), L["wkv"]).view(B, n + Q, 2, H, d)
            ops.attention(q.view(B, Q, H, d).permute(0, 2, 1, 3), kv[:, :, 0].permute(0, 2, 1, 3), kv[:, :, 1].permute(0, 2, 1, 3),
                          o.view(B, Q, H, d).permute(0, 2, 1, 3), scale=d ** -0.5)
            ops.gemm(o, L["wo"], out=lat, residual=lat)
            ops.layernorm(lat, L["nf"][0], L["nf"][1], 1e-5, out=lbuf)
            hmid = ops.gemm(lbuf, L["w1"], act=ops.ACT_GELU)
            ops.gemm(hmid, L["w2"], out=lat, residual=lat)
        hid = ops.layernorm(lat, self.norm_out[0], self.norm_out[1], 1e-5)                              # fp16 [B*Q, D]
        prompt = ops.gemm(hid, self.w_out, bias=self.b_out, out_dtype=torch.float32).view(B, Q, -1)
        # AttentionPool2d: token 0 = mean token; only its output is returned (resampler.py:89-116)
        t = torch.empty((B, Q + 1, D), device=dev, dtype=torch.float16)
        mean = ops.avgpool_tokens(hid.view(B, Q, D), Q)                                                 # [B,1,D]
        for b in range(B):
            ops.unary_f16(mean[b], out=t[b, :1])
            ops.unary_f16(hid[b * Q:(b + 1) * Q], out=t[b, 1:])
        t = ops.add_bcast_f16(t, self.pool_pos).view(B, Q + 1, D)
        P = self.pool
        pq = ops.gemm(t[:, 0], P["wq"], bias=P["bq"])                                                   # [B, D]
        pk = ops.gemm(t.view(B * (Q + 1), D), P["wk"], bias=P["bk"]).view(B, Q + 1, H, D // H)
        pv = ops.gemm(t.view(B * (Q + 1), D), P["wv"], bias=P["bv"]).view(B, Q + 1, H, D // H)
        po = torch.empty((B, D), device=dev, dtype=torch.float16)
        ops.attention(pq.view(B, 1, H, D // H).permute(0, 2, 1, 3), pk.permute(0, 2, 1, 3), pv.permute(0, 2, 1, 3),
                      po.view(B, 1, H, D // H).permute(0, 2, 1, 3), scale=(D // H) ** -0.5)
        pooled = ops.gemm(po, P["wc"], bias=P["bc"], out_dtype=torch.float32)
        return prompt, pooled

    forward = __call__

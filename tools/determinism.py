"""First diverging library call between two identical runs (allocator blocks churned in between)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth, ops
from seedx_b200.resampler_xl import ResamplerXLV2
from seedx_b200.sdxl import AutoencoderKL, EulerDiscreteScheduler, UNet2DConditionModel

log = []
def digest(t):
    t = t.contiguous()
    if t.dtype == torch.float16: v = t.view(torch.int16)
    elif t.dtype == torch.float32: v = t.view(torch.int32)
    else: v = t
    return int(v.to(torch.int64).sum().item()) ^ int((v.to(torch.int64) * torch.arange(v.numel(), device=v.device).view(v.shape) % 1000003).sum().item())
def wrap(name):
    f = getattr(ops, name)
    def g(*a, **k):
        r = f(*a, **k)
        outs = r if isinstance(r, (tuple, list)) else [r]
        desc = " ".join(str(tuple(x.shape)) for x in a if torch.is_tensor(x))
        log.append((name, desc, tuple(digest(o) for o in outs if torch.is_tensor(o))))
        return r
    setattr(ops, name, g)
for n in ["gemm", "conv2d_nhwc", "attention", "layernorm", "groupnorm_nhwc", "unary_f16", "avgpool_tokens", "add_bcast_f16", "im2col_nhwc",
          "upsample2x_nhwc", "timestep_embedding", "softmax_rows", "nchw_to_nhwc_f16", "nhwc_to_nchw_f32"]:
    if hasattr(ops, n): wrap(n)

def churn():
    xs = [torch.full((1 << 20,), float(i + 3), device="cuda") for i in range(64)]
    del xs
def compare(tag, f):
    global log
    log = []; f(); a = log
    churn()
    log = []; f(); b = log
    assert len(a) == len(b)
    bad = [i for i in range(len(a)) if a[i] != b[i]]
    print(f"== {tag}: {len(a)} calls, {len(bad)} differ")
    for i in bad[:6]:
        print("   call", i, a[i][0], a[i][1])

rcfg = dict(synth.TINY_RESAMPLER_XL, embedding_dim=256)
rx = ResamplerXLV2(normalize=False, **rcfg); rx.load_state_dict(synth.resampler_xl_state_dict(rcfg))
feats = synth.randn("adapter_feats", (1, 64, 256)).cuda()
compare("resampler_xl B=1", lambda: rx(feats))
f2 = torch.cat([feats, feats * 0.5])
compare("resampler_xl B=2", lambda: rx(f2))
ucfg = dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160)
unet = UNet2DConditionModel(ucfg); unet.load_state_dict(synth.unet_state_dict(ucfg))
p, pooled = rx(f2)
tid = torch.tensor([[256., 256., 0., 0., 256., 256.]], device="cuda").repeat(2, 1)
cond = unet.prepare_cond(p, pooled, tid)
xin = synth.randn("x", (2, 32, 32, 8)).cuda().half()
t = torch.full((2,), 981.0, device="cuda")
compare("unet tiny", lambda: unet.forward_nhwc(xin, t, cond))
vae = AutoencoderKL(synth.TINY_VAE); vae.load_state_dict(synth.vae_state_dict(synth.TINY_VAE))
lat = synth.randn("lat", (1, 4, 32, 32)).cuda()
compare("vae decode", lambda: vae.decode(lat))

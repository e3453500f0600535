"""Parity of the tcgen05 GEMM / implicit-GEMM conv (seedx_gemm_f16) against a torch fp32 reference of the same op.

Tolerance: operands are fp16 (exactly representable in fp32), accumulation fp32 -> the only differences are the
summation order and the final rounding to the output dtype: rel-Frobenius <= 2e-3 for fp16 out, 1e-5 for fp32 out.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


@pytest.fixture(params=["single_cta", "cluster2"])
def cluster_mode(request):
    """run each GEMM test with the 2-CTA cluster / TMA-multicast variant forced off and forced on"""
    from seedx_b200._lib import lib
    lib().seedx_gemm_set_cluster(0 if request.param == "single_cta" else 2)
    yield request.param
    lib().seedx_gemm_set_cluster(1)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (1024, 1664, 1664), (200, 328, 72),
                                   (2048, 4992, 1664), (174, 5120, 5120), (77, 40, 8), (1, 32330, 256)])
@pytest.mark.parametrize("tile_n", [0, 64, 96, 128, 144, 160, 192, 208, 224, 240, 256])
def test_gemm_plain(M, N, K, tile_n, cluster_mode):
    from seedx_b200 import ops
    a = mk((M, K), 1).half()
    w = mk((N, K), 2, K ** -0.5).half()
    ref = a.float() @ w.float().t()
    out32 = ops.gemm(a, w, out_dtype=torch.float32, tile_n=tile_n)
    assert rel(out32, ref) < 1e-5
    out16 = ops.gemm(a, w, out_dtype=torch.float16, tile_n=tile_n)
    assert rel(out16, ref) < 2e-3


@pytest.mark.parametrize("tile_n", [0, 144, 208])
def test_gemm_epilogues(tile_n, cluster_mode):
    from seedx_b200 import ops
    import functools
    M, N, K = 384, 768, 320
    ops = type("O", (), {k: getattr(ops, k) for k in dir(ops)})
    ops.gemm = functools.partial(ops.gemm, tile_n=tile_n)
    a = mk((M, K), 3).half()
    w = mk((N, K), 4, K ** -0.5).half()
    bias = mk((N,), 5)
    bias_m = mk((M,), 6)
    res32 = mk((M, N), 7)
    res16 = mk((M, N), 8).half()
    acc = a.float() @ w.float().t()
    # bias + gelu
    o = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, out_dtype=torch.float32)
    assert rel(o, F.gelu(acc + bias)) < 1e-5
    # bias + fp32 residual in place
    r = res32.clone()
    ops.gemm(a, w, out=r, bias=bias, residual=r)
    assert rel(r, acc + bias + res32) < 1e-5
    # fp16 residual, fp16 out, alpha, bias_m
    o = ops.gemm(a, w, bias_m=bias_m, residual=res16, alpha=0.5, out_dtype=torch.float16)
    assert rel(o, 0.5 * acc + bias_m[:, None] + res16.float()) < 2e-3
    # gated (GEGLU with interleaved columns): y[j] = x[2j] * gelu(x[2j+1])
    o = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, gated=True, out_dtype=torch.float32)
    x = acc + bias
    assert rel(o, x[:, 0::2] * F.gelu(x[:, 1::2])) < 1e-5
    o = ops.gemm(a, w, act=ops.ACT_SILU, gated=True, out_dtype=torch.float16)
    assert rel(o, acc[:, 0::2] * F.silu(acc[:, 1::2])) < 2e-3
    # row-modulo residual (positional-embedding broadcast) + per-row-group bias
    pos = mk((128, N), 9)
    bg = mk((3, N), 10)
    o = ops.gemm(a, w, residual=pos, res_row_mod=128, bias_g=bg, bias_g_rows=128, out_dtype=torch.float32)
    assert rel(o, acc + pos.repeat(3, 1) + bg.repeat_interleave(128, 0)) < 1e-5


def test_gemm_batched_strided(cluster_mode):
    from seedx_b200 import ops
    B, M, N, K = 5, 200, 136, 104
    a_full = mk((B, M, 3 * K + 8), 11).half()
    a = a_full[:, :, K:2 * K]           # strided view (lda = 3K+8), offset 16B aligned (K*2 = 208 B)
    w = mk((B, N, K), 12, K ** -0.5).half()
    ref = torch.einsum("bmk,bnk->bmn", a.float(), w.float())
    o = ops.gemm(a, w, out_dtype=torch.float32)
    assert rel(o, ref) < 1e-5
    w1 = w[0].contiguous()
    o = ops.gemm(a, w1, out_dtype=torch.float32)  # shared B
    assert rel(o, torch.einsum("bmk,nk->bmn", a.float(), w1.float())) < 1e-5


@pytest.mark.parametrize("n,h,w,c,cout,taps", [(2, 32, 32, 128, 192, 3), (1, 64, 64, 320, 320, 3), (2, 128, 128, 64, 64, 3),
                                               (1, 256, 256, 128, 24, 3), (3, 16, 16, 8, 320, 3), (2, 32, 32, 960, 640, 1),
                                               (1, 128, 128, 8, 320, 3)])
def test_conv_nhwc(n, h, w, c, cout, taps, cluster_mode):
    from seedx_b200 import ops
    x = mk((n, c, h, w), 21).half()
    wt = mk((cout, c, taps, taps), 22, (c * taps * taps) ** -0.5).half()
    bias = mk((cout,), 23)
    temb = mk((n, cout), 24)
    ref = F.conv2d(x.float(), wt.float(), bias=bias, padding=taps // 2) + temb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).contiguous()
    cpad = (c + 63) // 64 * 64
    wp = torch.zeros((cout, taps, taps, cpad), dtype=torch.float16, device="cuda")
    wp[..., :c] = wt.permute(0, 2, 3, 1)
    wp = wp.reshape(cout, taps * taps * cpad).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous()
    o = ops.conv2d_nhwc(xn, wp, taps=taps, bias=bias, bias_g=temb, out_dtype=torch.float32)
    assert rel(o, ref) < 1e-5
    res = mk((n, h, w, cout), 25).half()
    o = ops.conv2d_nhwc(xn, wp, taps=taps, bias=bias, bias_g=temb, residual=res, out_dtype=torch.float16)
    assert rel(o, ref + res.float()) < 2e-3


# ---- the bench's hot shapes: long-K 3x3 convs with C >= 640 in pair mode with the TMA-residual epilogue, and the
# M = 8192 x N = 10240 gated GEGLU projection, each with the tile the auto picker chooses --------------------------------------------------
@pytest.mark.parametrize("n,h,w,c,cout", [(2, 32, 32, 1280, 1280), (2, 64, 64, 640, 640), (1, 32, 32, 2560, 1280), (1, 64, 64, 1920, 640),
                                          (1, 128, 128, 960, 320)])
def test_conv_nhwc_unet_hot_shapes(n, h, w, c, cout):
    """3x3 convs of the SDXL UNet at their real widths (K = 9*C up to 23 040: >= 90 k-blocks through the stage ring, CTA pairs, bias + per-image
    time-embedding add + fp16 residual through the TMA epilogue), fp32 torch reference on the device"""
    from seedx_b200 import ops
    x = mk((n, c, h, w), 31).half()
    wt = mk((cout, c, 3, 3), 32, (c * 9) ** -0.5).half()
    bias, temb = mk((cout,), 33), mk((n, cout), 34)
    ref = (F.conv2d(x.float(), wt.float(), bias=bias, padding=1) + temb[:, :, None, None]).permute(0, 2, 3, 1).contiguous()
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * c).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous()
    res = mk((n, h, w, cout), 35).half()
    o = ops.conv2d_nhwc(xn, wp, taps=3, bias=bias, bias_g=temb, residual=res, out_dtype=torch.float16)
    assert rel(o, ref + res.float()) < 2e-3
    o32 = ops.conv2d_nhwc(xn, wp, taps=3, bias=bias, bias_g=temb, out_dtype=torch.float32)
    assert rel(o32, ref) < 3e-5          # fp32 accumulation over up to 23 040 products: the summation order alone moves this to ~1e-5


@pytest.mark.parametrize("M,N,K,kind", [(8192, 10240, 1280, "geglu"), (32768, 5120, 640, "geglu"), (8192, 1280, 5120, "residual"),
                                        (8192, 1280, 1280, "residual"), (32768, 640, 2560, "residual"), (8192, 3840, 1280, "plain")])
def test_gemm_unet_hot_shapes(M, N, K, kind):
    """the transformer GEMMs of the UNet forward at bench size (8 samples): GEGLU projection with bias + gating, attention-out / feed-forward-down
    with bias + in-place fp16 residual, fused QKV"""
    from seedx_b200 import ops
    a = mk((M, K), 41).half()
    w = mk((N, K), 42, K ** -0.5).half()
    bias = mk((N,), 43)
    acc = a.float() @ w.float().t() + bias
    if kind == "geglu":
        o = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, gated=True)
        assert rel(o, acc[:, 0::2] * F.gelu(acc[:, 1::2])) < 2e-3
    elif kind == "residual":
        r = mk((M, N), 44).half()
        want = acc + r.float()
        ops.gemm(a, w, out=r, bias=bias, residual=r)
        assert rel(r, want) < 2e-3
    else:
        assert rel(ops.gemm(a, w, bias=bias), acc) < 2e-3


@pytest.mark.parametrize("M,N,K", [(8192, 3840, 1280), (4096, 640, 640), (300, 1280, 1280), (1024, 10240, 1280)])
@pytest.mark.parametrize("mean_shift", [0.0, 3.0])
def test_gemm_folded_layernorm(M, N, K, mean_shift, cluster_mode):
    """LayerNorm folded into the projection (seedx_gemm_args.ln_stats): rows of x are the A operand un-normalised, B = W * gamma, the epilogue
    applies rstd * (acc - mean * colsum) + (W . beta + b).  Reference = fp32 F.layer_norm followed by the linear layer; rows with a mean several
    sigma away from zero check the cancellation in acc - mean * colsum."""
    from seedx_b200 import ops
    x = (mk((M, K), 51) * (1.0 + mk((M, 1), 52).abs()) + mean_shift * mk((M, 1), 53)).half()
    w = mk((N, K), 54, K ** -0.5)
    gamma, beta, b = 1.0 + 0.2 * mk((K,), 55), 0.2 * mk((K,), 56), mk((N,), 57)
    ref = F.layer_norm(x.float(), (K,), gamma, beta, 1e-5) @ w.t() + b
    wf = (w * gamma[None, :]).half()
    colsum = wf.float().sum(1).contiguous()
    bias = (w @ beta + b).contiguous()
    st = ops.row_stats(x, 1e-5)
    xf = x.float()
    assert rel(st[:, 0], xf.mean(1)) < 1e-5 and rel(st[:, 1], torch.rsqrt(xf.var(1, unbiased=False) + 1e-5)) < 1e-5
    o = ops.gemm(x, wf, bias=bias, ln=(st, colsum), out_dtype=torch.float32)
    assert rel(o, ref) < 1e-3          # W * gamma is rounded to fp16 once (the reference rounds the normalised activations instead)
    if N % 2 == 0 and N >= 640:
        og = ops.gemm(x, wf, bias=bias, ln=(st, colsum), act=ops.ACT_GELU, gated=True)
        assert rel(og, ref[:, 0::2] * F.gelu(ref[:, 1::2])) < 3e-3


@pytest.mark.parametrize("M,N,K,res", [(8192, 1280, 1280, True), (4096, 640, 2560, True), (1024, 320, 64, False), (256, 96, 128, True)])
def test_gemm_epilogue_statistics(M, N, K, res, cluster_mode):
    """row_part / col_part (seedx_gemm_args): (sum, sum of squares) of the stored output per row over each 32-column chunk and per column over each
    32-row slab, then the consumers: a GEMM with the LayerNorm folded in that forms mean / rstd from the row partials, GroupNorm from the column partials"""
    from seedx_b200 import ops
    a = mk((M, K), 61).half()
    w = mk((N, K), 62, K ** -0.5).half()
    bias = mk((N,), 63)
    r = mk((M, N), 64).half() if res else None
    want = a.float() @ w.float().t() + bias + (r.float() if res else 0)
    rp = torch.empty((N // 32, M, 2), device="cuda")
    cp = torch.empty((M // 32, N, 2), device="cuda")
    out = r.clone() if res else None
    st_epi = torch.empty((M, 2), device="cuda")
    tickets = torch.zeros((M // 32,), device="cuda", dtype=torch.int32)
    out = ops.gemm(a, w, out=out, bias=bias, residual=out if res else None, row_part=rp, col_part=cp, row_stats=(st_epi, tickets, 1e-5))
    assert rel(out, want) < 2e-3
    assert int(tickets.abs().sum()) == 0                                # the slab tickets reset themselves
    assert rel(st_epi, ops.row_finalize(rp, 1e-5)) < 1e-5              # finalized in the epilogue (chunk order) = the finalize kernel (4 interleaved sub-sums)
    st2 = torch.empty_like(st_epi)
    ops.gemm(a, w, out=(r.clone() if res else None), bias=bias, residual=None, row_part=rp, row_stats=(st2, tickets, 1e-5)) if not res else None
    if not res:
        assert torch.equal(st2, st_epi)                                 # run-to-run bit-reproducible whichever warp comes last
    o32 = out.float()
    v = want.view(M, N // 32, 32)
    assert rel(rp[..., 0].t(), v.sum(-1)) < 2e-3 and rel(rp[..., 1].t(), (v * v).sum(-1)) < 2e-3
    fin = ops.row_finalize(rp, 1e-5)           # (mean, rstd) from the partials = the statistics of the stored rows
    assert rel(fin[:, 0], o32.mean(1)) < 2e-3 and rel(fin[:, 1], torch.rsqrt(o32.var(1, unbiased=False) + 1e-5)) < 2e-3
    c = o32.view(M // 32, 32, N)              # column partials are taken from the stored fp16 values: exact up to the summation order
    assert rel(cp[..., 0], c.sum(1)) < 1e-5 and rel(cp[..., 1], (c * c).sum(1)) < 1e-5
    if N % 64 == 0:
        # consumer 1: LayerNorm(out) -> linear, statistics from the row partials
        w2 = mk((256, N), 65, N ** -0.5)
        gamma, beta = 1.0 + 0.2 * mk((N,), 66), 0.2 * mk((N,), 67)
        ref = F.layer_norm(o32, (N,), gamma, beta, 1e-5) @ w2.t()
        wf = (w2 * gamma[None, :]).half()
        o2 = ops.gemm(out, wf, bias=(w2 @ beta).contiguous(), ln=(rp, wf.float().sum(1).contiguous(), 1e-5), out_dtype=torch.float32)
        assert rel(o2, ref) < 1.5e-3


@pytest.fixture(params=["data_parallel", "stream_k"])
def stream_k_mode(request):
    """run a test with the stream-K schedule forced off and forced on wherever it is legal"""
    from seedx_b200._lib import lib
    lib().seedx_gemm_set_stream_k(0 if request.param == "data_parallel" else 2)
    yield request.param
    lib().seedx_gemm_set_stream_k(1)


@pytest.mark.parametrize("M,N,K,tile_n", [(8192, 1280, 1280, 0), (8192, 1280, 5120, 224), (8192, 1280, 320, 256), (32768, 640, 640, 0), (4096, 3072, 1024, 160),
                                          (2500, 1000, 4096, 128), (19000, 520, 192, 64), (8192, 10240, 1280, 0)])
def test_gemm_stream_k(M, N, K, tile_n, stream_k_mode, cluster_mode):
    """stream-K: clusters take equal shares of the (tile, k-block) iterations; tiles cut between clusters are completed through the fp32 partial
    workspace in a fixed order.  Same results as the data-parallel schedule to fp32 summation order, bit-identical from run to run, with every
    epilogue kind (bias, gating, in-place residual through the TMA epilogue, fp32 output)."""
    from seedx_b200 import ops
    a = mk((M, K), 71).half()
    w = mk((N, K), 72, K ** -0.5).half()
    bias = mk((N,), 73)
    acc = a.float() @ w.float().t() + bias
    o32 = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, tile_n=tile_n)
    assert rel(o32, acc) < 1e-5
    assert torch.equal(o32, ops.gemm(a, w, bias=bias, out_dtype=torch.float32, tile_n=tile_n))
    r = mk((M, N), 74).half()
    want = acc + r.float()
    ops.gemm(a, w, out=r, bias=bias, residual=r, tile_n=tile_n)
    assert rel(r, want) < 2e-3
    if N % 2 == 0:
        o = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, gated=True, tile_n=tile_n)
        assert rel(o, acc[:, 0::2] * F.gelu(acc[:, 1::2])) < 2e-3


@pytest.mark.parametrize("n,h,w,c,cout", [(8, 32, 32, 1280, 1280), (2, 64, 64, 640, 640), (3, 32, 32, 320, 640)])
def test_conv_stream_k(n, h, w, c, cout, stream_k_mode):
    from seedx_b200 import ops
    x = mk((n, c, h, w), 81).half()
    wt = mk((cout, c, 3, 3), 82, (c * 9) ** -0.5).half()
    bias, temb = mk((cout,), 83), mk((n, cout), 84)
    ref = (F.conv2d(x.float(), wt.float(), bias=bias, padding=1) + temb[:, :, None, None]).permute(0, 2, 3, 1).contiguous()
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * c).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous()
    res = mk((n, h, w, cout), 85).half()
    part = torch.empty((n * h * w // 32, cout, 2), device="cuda")
    o = ops.conv2d_nhwc(xn, wp, taps=3, bias=bias, bias_g=temb, residual=res, col_part=part)
    assert rel(o, ref + res.float()) < 2e-3
    c32 = o.float().view(-1, 32, cout)
    assert rel(part[..., 0], c32.sum(1)) < 1e-5

#!/usr/bin/env python
"""bench.py — end-to-end images/sec of the SEED-X hot path (448x448 in -> 1024x1024 out) on N B200s of one node.

Step = one pass of the whole pipe over one batch of B synthetic requests per GPU:
  448^2 image (any-res '1x1' grid -> 2 ViT views) + 32-token prompt ending in <img>
  -> ViT-bigG -> LLaMA-13B prefill + 66 greedy tokens (65 of them forced by the image-token logits processor) -> output resampler
  -> ResamplerXLV2 -> 50 Euler steps x 2-way CFG SDXL UNet @128^2 latents -> VAE decode -> 1024^2 uint8 image.
Weights: random-init at the full architecture sizes (no checkpoints offline).  One process per GPU; ranks are independent
replicas (weak scaling), NCCL only gathers the finished images.

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = pinned-host inputs copied in and images copied out
inside the timed region.  `roofline` is for the dominant launch, the CUDA-graphed UNet sample-forward (6.75 TFLOP per
sample); `cpu_baseline` times the CPU oracle (oracle/vit.py, restatement of the reference ViT) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# analytic work per image, TFLOP (SURVEY.md Appendix A.4 / BASELINE.md §3)
VIT_TFLOP_448, VIT_TFLOP_224 = 4.219, 1.011
UNET_TFLOP, VAE_DEC_TFLOP = 6.75, 10.47
LLM_GFLOP_TOK = 25.7
METRIC = "end-to-end images/sec (448x448 in -> 1024x1024 out)"


def work_per_image_tflop(p_len=175, new_tokens=66, steps=50):
    llm = (p_len + new_tokens) * LLM_GFLOP_TOK / 1e3
    return 2 * VIT_TFLOP_448 + llm + 2 * steps * UNET_TFLOP + VAE_DEC_TFLOP


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop = gpu_index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=3)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def pick_threads():
    """Thread count for the CPU arm = the one that is actually fastest on this host: more threads than the container's real core budget
    (cgroup quota, SMT siblings) make torch's CPU GEMMs slower, not faster, so 'all the threads it can use' is calibrated on a 2048^3 fp32
    matmul over a ladder of candidates instead of taken from os.cpu_count()."""
    import torch
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({c for c in (4, 8, 16, 32, 64, 128, 256, avail) if 1 <= c <= avail}) or [1]
    a, b = torch.randn(2048, 2048), torch.randn(2048, 2048)
    best_c, best_rate = cands[0], 0.0
    for c in cands:
        torch.set_num_threads(c)
        a @ b
        t0 = time.time()
        for _ in range(2):
            a @ b
        rate = 2 * 2 * 2048 ** 3 / max(time.time() - t0, 1e-9)
        if rate > best_rate * 1.03:
            best_c, best_rate = c, rate
    torch.set_num_threads(best_c)
    return best_c, avail


def cpu_vit_sample(layers, threads, with_output=False):
    """time the CPU oracle (reference ViT restated, fp32) on config 1: one 224x224 image, `layers` of the 48 blocks + pool."""
    import torch
    from oracle import vit as ovit
    from seedx_b200 import synth
    torch.set_num_threads(threads)
    cfg = dict(width=1664, layers=layers, heads=16, mlp_width=8192, output_dim=4096, n_queries=256, patch=14)
    sd = synth.vit_state_dict(**cfg)
    x = synth.image("bench_cpu_img", 1, 224)
    t0 = time.time()
    out = ovit.vit_forward(sd, x, 16)
    dt = time.time() - t0
    tflop = (VIT_TFLOP_224 - 0.0476) * layers / 48 + 0.0476          # blocks scale with depth; pool/head ~0.048 TFLOP
    return dt, tflop, (sd, x, out) if with_output else None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port of its PyTorch modules), host cores."""
    cores, avail = pick_threads()
    layers = 12
    times = []
    for i in range(args.warmup + args.steps):
        dt, tflop, _ = cpu_vit_sample(layers, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = statistics.mean(times) * 1e3
    cpu_tflops = tflop / (ms / 1e3)
    ips = cpu_tflops / work_per_image_tflop()
    sample = f"oracle ViT-bigG fp32, one 224x224 image, {layers}/48 blocks + attention pool ({tflop:.3f} TFLOP) per step on {cores} threads (fastest of the " \
             f"ladder up to the {avail} schedulable CPUs); images/s extrapolated by FLOPs to the full pipe ({work_per_image_tflop():.0f} TFLOP/image)"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "seedx_i2i_448_to_1024 (CPU sample, see cpu_baseline.sample)"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample, "cpu_tflops": cpu_tflops},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_comprehension(args, eng, rank, world, peaks_):
    """--workload comprehension = BASELINE.json configs[1]: 1 x 448^2 image (any-res 1x1 -> 2 ViT views) + 32-token prompt -> 128 greedy text tokens,
    batch 1 per GPU (EOS suppressed so every step decodes exactly 128 tokens).  Auxiliary line: tokens/s, ms per token, decode HBM roofline."""
    import torch
    from seedx_b200 import _lib, synth, trace
    from seedx_b200 import dist as sdist
    tensor_peak, hbm_peak, peak_src = peaks_
    new_tok = 128
    views_host = synth.image(f"bench_cmp_views_rank{rank}", 2, 448).pin_memory()
    views_dev = views_host.cuda()
    patch_pos = torch.tensor([[0.5, 0.5], [0.5, 0.5]])
    text_ids = torch.randint(3, eng.tok.base, (32,), generator=torch.Generator().manual_seed(4321 + rank)).tolist()
    ids, mask = eng.build_prompt(2, text_ids, force_image=False)
    P = int(ids.numel())

    def step(e2e):
        v = views_host.cuda(non_blocking=True) if e2e else views_dev
        feats = eng.vit(v)
        trace.mark("vit")
        req = dict(input_ids=ids.unsqueeze(0), image_embeds=feats, embeds_cmp_mask=torch.ones((2, 64), dtype=torch.bool), ids_cmp_mask=mask.unsqueeze(0),
                   patch_positions=patch_pos)
        out = eng.agent.generate_batch(eng.tok, [req], max_new_tokens=new_tok, suppress_eos=True)[0]   # reads the ids back to the host (d2h)
        return out

    def timed(e2e):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(args.steps):
            out = step(e2e)
        e1.record()
        torch.cuda.synchronize()
        assert len(out["ids"]) == new_tok
        return sdist.max_over_ranks(e0.elapsed_time(e1), "cuda"), _lib.launch_count() - n0

    for _ in range(args.warmup):
        step(False)
    ms_dev, launches = timed(False)
    ms_e2e, _ = timed(True)
    trace.enable(True)
    trace.mark("start")
    step(False)
    detail = trace.summary()
    trace.enable(False)
    if rank == 0:
        toks = world * new_tok * args.steps
        dec_ms = detail.get("llm.decode", 0.0)
        gb = 26.04 * (new_tok - 1)
        print(json.dumps({
            "metric": "comprehension text tokens/sec (1x448^2 image + 32-token prompt -> 128 greedy tokens, batch 1 per GPU)", "value": toks / (ms_dev / 1e3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "seedx_comprehension_448 (BASELINE.json configs[1]): prompt %d tokens incl. 2 x 64 image rows, 128 new tokens, EOS suppressed" % P,
                       "weights": "random-init, full sizes", "l2": "26 GB of weights per token step exceed the 126 MB L2"},
            "stage_detail_ms": detail,
            "e2e": {"value": toks / (ms_e2e / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": views_host.numel() * 4 + P * 8,
                    "d2h_bytes_per_step": (P + new_tok) * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "CUDA-graph replay of one decode step (gemv_mma_kernel = 84 % of it)", "achieved": gb / (dec_ms / 1e3) if dec_ms else None,
                         "peak": hbm_peak, "unit": "GB/s", "frac": gb / (dec_ms / 1e3) / hbm_peak if dec_ms else None, "traffic": None,
                         "ms_per_token": dec_ms / (new_tok - 1), "bytes_per_token_gb": 26.04, "peak_source": peak_src}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SEEDX_BENCH_BATCH", "4")), help="requests per GPU per step")
    ap.add_argument("--denoise-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small", action="store_true", help="debug: tiny models (NOT a valid benchmark)")
    ap.add_argument("--workload", default="i2i", choices=["i2i", "comprehension"],
                    help="i2i (default, the headline metric: 448^2 in -> 1024^2 out) | comprehension (BASELINE.json configs[1], auxiliary tokens/s line)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from seedx_b200 import _lib, synth
    from seedx_b200 import dist as sdist
    from seedx_b200.engine import SeedXEngine
    log = (lambda *a: print("[bench]", *a, file=sys.stderr, flush=True)) if rank == 0 else None
    tensor_peak, hbm_peak, peak_src = peaks()

    # ---- CPU baseline first (rank 0, N=1): also yields CPU-generated ViT weights for a full-depth parity check -------------
    cpu_base, vit_sd, parity = None, None, None
    do_cpu = (world == 1 and rank == 0 and not args.no_cpu_baseline and not args.small)
    if do_cpu:
        cores, avail = pick_threads()
        dt, tflop, (vit_sd, x224, ref224) = cpu_vit_sample(48, cores, with_output=True)
        cpu_tflops = tflop / dt
        cpu_base = {"value": cpu_tflops / work_per_image_tflop(), "unit": "images/s", "cores": cores, "kind": "port",
                    "sample": f"oracle (reference ViT restated, fp32 torch, {cores} threads = fastest of the ladder up to the {avail} schedulable CPUs): config 1 in full = one 224x224 image through ViT-bigG "
                              f"({tflop:.2f} TFLOP) in {dt:.1f} s = {cpu_tflops:.3f} TFLOP/s; images/s extrapolated by FLOPs to the full pipe "
                              f"({work_per_image_tflop():.0f} TFLOP/image)", "cpu_tflops": cpu_tflops, "sample_seconds": dt}
        log(f"cpu baseline: {dt:.1f}s, {cpu_tflops:.3f} TFLOP/s on {cores} threads")

    if args.small:
        eng = SeedXEngine(vit_cfg=dict(width=208, layers=2, heads=2, mlp_width=520, output_dim=256, n_queries=256, patch=14),
                          llm_cfg=synth.TINY_LLAMA, unet_cfg=dict(synth.TINY_UNET, cross_attention_dim=256, text_embed_dim=160),
                          vae_cfg=synth.TINY_VAE, rxl_cfg=dict(synth.TINY_RESAMPLER_XL, embedding_dim=256, output1_dim=96, output2_dim=160),
                          log=log)
    else:
        eng = SeedXEngine(vit_sd=vit_sd, log=log)
    if do_cpu:
        out224 = eng.vit(x224.cuda())
        parity = {"vit_224_full_depth_rel_err_vs_oracle": float(((out224.float().cpu() - ref224).norm() / ref224.norm()).item()), "tolerance": 1e-3}
        log(f"full-depth ViT parity vs oracle: {parity}")
        del vit_sd

    if args.workload == "comprehension":
        run_comprehension(args, eng, rank, world, (tensor_peak, hbm_peak, peak_src))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    B = args.batch
    n_views = 2
    # ---- synthetic inputs: post-transform views in pinned host memory + device-resident copies ------------------------------
    views_host = synth.image(f"bench_views_rank{rank}", B * n_views, 448).pin_memory()
    views_dev = views_host.cuda()
    patch_pos = torch.tensor([[0.5, 0.5], [0.5, 0.5]] * B)
    g = torch.Generator().manual_seed(1234 + rank)
    text_ids = [torch.randint(3, eng.tok.base, (32,), generator=g).tolist() for _ in range(B)]
    out_host = torch.empty((B, 1024, 1024, 3), dtype=torch.uint8).pin_memory() if not args.small else None
    h2d_bytes = views_host.numel() * 4 + B * 32 * 8
    gather_buf = None

    def step(e2e):
        v = views_host.cuda(non_blocking=True) if e2e else views_dev
        u8 = eng.generate(v, patch_pos, text_ids, steps=args.denoise_steps, n_views=n_views)
        nonlocal gather_buf
        if world > 1:                                  # the only collective: gather finished images on every rank (NCCL / NVLink)
            gather_buf = sdist.gather_images(u8, gather_buf)
        if e2e:
            if out_host is not None:
                out_host.copy_(u8, non_blocking=True)
            else:
                u8.cpu()
        return u8

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(e2e=(i % 2 == 1))
        if log:
            log(f"warmup {i}: {eng.stage_ms()}")
    sync()

    def timed(e2e):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        stages = []
        e0.record()
        for _ in range(args.steps):
            step(e2e)
            stages.append(eng._events)
        e1.record()
        sync()
        ms = sdist.max_over_ranks(e0.elapsed_time(e1), "cuda")
        st = {k: statistics.mean(a.elapsed_time(b) for (a, b) in [(ev[i], ev[i + 1]) for ev in stages])
              for i, k in enumerate(("vit_ms", "llm_ms", "detok_ms"))}
        return ms, _lib.launch_count() - n0, st

    if rank == 0:       # one nvidia-smi poller per job (8 of them would only add host noise); rank 0's GPU stands for the box
        with ClockSampler(local) as cs:
            ms_dev, launches, stages = timed(e2e=False)
            ms_e2e, _, _ = timed(e2e=True)
        clocks = cs.summary()
    else:
        ms_dev, launches, stages = timed(e2e=False)
        ms_e2e, _, _ = timed(e2e=True)
        clocks = None
    # one extra untimed step with section marks on (seedx_b200.trace): where the step goes, stage by stage
    from seedx_b200 import trace
    trace.enable(True)
    step(e2e=False)
    detail = trace.summary()
    # decode probe: the bench prompt ends in <img>, so its 65 forced image tokens ride through the prefill pass (jump-forward) and the step
    # has almost no token loop left; the HBM-bound decode path is therefore timed separately on 64 free-running token steps of B lock-step
    # sequences (64-token random prompts, EOS suppressed) — untimed for the headline, reported in stage_roofline.llm_decode
    decode_step_ms = None
    if not args.small:
        trace.enable(True)
        pg = torch.Generator().manual_seed(99 + rank)
        p_ids = [torch.randint(3, eng.tok.base, (64,), generator=pg) for _ in range(B)]
        p_emb = [eng.llm.get_input_embeddings()(i)[0] for i in p_ids]
        eng.llm.generate_greedy_batch(p_ids, p_emb, img_ids=None, max_new_tokens=65, eos_id=None, suppress_eos=True)
        decode_step_ms = trace.summary().get("llm.decode", 0.0) / 64
    trace.enable(False)
    # transparency: the same timed loop with jump-forward off (every forced image token takes its own decode step, as in the reference)
    token_loop = None
    if world == 1 and not args.small and eng.llm.jump_forward:
        try:
            eng.llm.jump_forward = False
            step(e2e=False)
            ms_tl, _, st_tl = timed(e2e=False)
            token_loop = {"value": B * args.steps / (ms_tl / 1e3), "unit": "images/s", "ms_per_step": ms_tl / args.steps, "stage_ms_per_step": st_tl,
                          "note": "SEEDX_JUMP_FORWARD=0 behaviour: the 65 forced image tokens decoded one step at a time"}
        except Exception as exc:                          # never let the extra measurement take the headline down
            token_loop = {"error": repr(exc)}
        finally:
            eng.llm.jump_forward = True

    if rank == 0:
        imgs = world * B * args.steps
        value = imgs / (ms_dev / 1e3)
        e2e_v = imgs / (ms_e2e / 1e3)
        # dominant launch: the CUDA-graphed UNet sample-forward (2B samples per launch, 50 launches per step)
        # its time = de-tokenizer stage minus (resampler + VAE), measured live below on the same stream
        loop = next(iter(eng.adapter._loops.values()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            loop.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        unet_ms = e0.elapsed_time(e1) / reps
        achieved = UNET_TFLOP * 2 * B / (unet_ms / 1e3) if not args.small else None
        # the dominant kernel's dominant shape, timed alone with CUDA events on the launching stream: gemm_tc_kernel<256,2> on the GEGLU
        # projection of the UNet feed-forward (60 launches per forward = 16 % of it, the largest single line of the launch list); three
        # operand sets (393 MB) rotate so that no launch finds its inputs in the 126 MB L2
        dom = None
        if not args.small:
            from seedx_b200 import ops
            Mg, Ng, Kg = 2 * B * 1024, 10240, 1280
            As = [torch.randn(Mg, Kg, device="cuda").half() for _ in range(3)]
            Ws = [(torch.randn(Ng, Kg, device="cuda") * 0.03).half() for _ in range(3)]
            Os = [torch.empty(Mg, Ng // 2, device="cuda", dtype=torch.float16) for _ in range(3)]
            gb = torch.randn(Ng, device="cuda")
            for j in range(3):
                ops.gemm(As[j], Ws[j], out=Os[j], bias=gb, act=ops.ACT_GELU, gated=True)
            torch.cuda.synchronize()
            n_l = 30
            e0.record()
            for j in range(n_l):
                ops.gemm(As[j % 3], Ws[j % 3], out=Os[j % 3], bias=gb, act=ops.ACT_GELU, gated=True)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n_l
            fl = 2.0 * Mg * Ng * Kg
            dom = {"kernel": "gemm_tc_kernel<256,2> (tcgen05 cta_group::2), GEGLU projection M=%d N=%d K=%d, bias + GELU gating fused" % (Mg, Ng, Kg),
                   "launch_us": us, "achieved": fl / us / 1e6, "unit": "TFLOP/s", "peak": tensor_peak, "frac": fl / us / 1e6 / tensor_peak,
                   "algorithmic_bytes": 2 * (Mg * Kg + Ng * Kg + Mg * Ng // 2),
                   "traffic": 95.8e6 if B == 4 else None,
                   "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of this shape, ncu --set full: profiles/r01_ncu_full_geglu_gemm_final.md"}
            del As, Ws, Os
        # per-stage achieved rate against the roofline that bounds the stage, from the marked extra step
        P = int(eng.build_prompt(n_views, text_ids[0])[0].numel())
        new_tok = 66

        def tens(tflop, ms):
            return {"bound": "tensor", "achieved": tflop / (ms / 1e3), "unit": "TFLOP/s", "peak": tensor_peak, "frac": tflop / (ms / 1e3) / tensor_peak,
                    "ms": ms}
        stage_roof = None
        if not args.small and detail:
            dec_ms = decode_step_ms * 64
            gb = 26.04 * 64
            stage_roof = {
                "vit": tens(B * n_views * VIT_TFLOP_448, detail["vit"]),
                "llm_prefill": dict(tens(B * (P + new_tok - 1) * LLM_GFLOP_TOK / 1e3, detail["llm.prefill"]), prompt_len=P, forced_rows=new_tok - 1,
                                    note="rows per request = prompt + the 65 forced image tokens (jump-forward); all requests of a step share one "
                                         "pass over the weights"),
                "llm_decode": {"bound": "hbm", "achieved": gb / (dec_ms / 1e3), "unit": "GB/s", "peak": hbm_peak, "frac": gb / (dec_ms / 1e3) / hbm_peak,
                               "ms": dec_ms, "ms_per_token_step": decode_step_ms, "sequences_in_lock_step": B, "bytes_per_step_gb": 26.04,
                               "note": "probe outside the timed step: 64 free-running token steps (the step's own forced image tokens are "
                                       "teacher-forced rows of the prefill pass); token loop left inside the step: %.1f ms" % detail.get("llm.decode", 0.0)},
                "unet_denoise_loop": tens(2 * B * args.denoise_steps * UNET_TFLOP, detail["detok.denoise_loop"]),
                "vae_decode": tens(B * VAE_DEC_TFLOP, detail["detok.vae_decode"]),
            }
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": "seedx_i2i_448_to_1024: per GPU %d requests/step, each 1x448^2 image (any-res 1x1 -> 2 ViT views) + 32-token "
                                   "prompt -> ViT-bigG -> LLaMA-13B prefill + 66 greedy tokens (65 forced by the image-token logits processor: teacher-forced rows of "
                                   "the prefill pass, results identical to token-by-token decoding; SEEDX_JUMP_FORWARD=0 restores the loop) -> ResamplerXLV2 -> "
                                   "%d Euler steps x 2-way CFG SDXL UNet -> VAE decode -> 1024^2 uint8" % (B, args.denoise_steps),
                       "requests_per_gpu": B, "denoise_steps": args.denoise_steps, "parallelism": f"replicas x{world} (one request set per rank)",
                       "l2": "working set (35 GB fp16 weights/GPU) exceeds the 126 MB L2; no explicit flush", "weights": "random-init, full sizes",
                       "small_debug_models": bool(args.small)},
            "stage_ms_per_step": stages,
            "stage_detail_ms": detail,
            "stage_roofline": stage_roof,
            "without_jump_forward": token_loop,
            "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * 1024 * 1024 * 3,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "CUDA-graph launch of one UNet sample-forward: 960 kernels, gemm_tc_kernel (GEMM + implicit-GEMM conv) = 71% of its device time, tcgen05 attention 21% (profiles/r01_unet_forward_launches_final.md)",
                         "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": (achieved / tensor_peak) if achieved else None,
                         "dominant_kernel": dom, "traffic": None, "traffic_note": "per-kernel DRAM bytes of the dominant GEMM shape: profiles/r01_ncu_full_geglu_gemm_final.md (95.8 MB/launch vs 131 MB algorithmic)", "peak_source": peak_src, "flop_per_launch": UNET_TFLOP * 2 * B * 1e12,
                         "launch_ms": unet_ms, "work_per_image_tflop": work_per_image_tflop(),
                         "pipeline_frac": (value / world) * work_per_image_tflop() / tensor_peak},
            "cpu_baseline": cpu_base,
            "parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()          # rank 0 is still measuring the roofline launches: nobody tears the communicator down before it is done
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

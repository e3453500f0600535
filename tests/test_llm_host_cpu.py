"""Host logic of the stage-2 engine on the CPU: `seedx_b200.llm.LlamaForCausalLM` runs unchanged, but its `ops` module is replaced by
tests/fake_ops.py (plain-torch restatements of the C entry points — a test double, see its header).  What is checked here is everything the
Python host decides: paged-KV bookkeeping, lock-step batching of ragged prompts, jump-forward over forced image spans, budget / EOS handling and
the HF-style generate surface — against the goldens produced by the reference's own modules (tests/golden/llama_tiny.pt)."""
import os

import pytest
import torch

import fake_ops
from seedx_b200 import llm as llm_mod
from seedx_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.fixture
def host(monkeypatch):
    monkeypatch.setattr(llm_mod, "ops", fake_ops)
    g = torch.load(os.path.join(GOLD, "llama_tiny.pt"))
    cfg = synth.TINY_LLAMA
    sd = synth.llama_state_dict(cfg)
    tok = synth.SynthTokenizer(vocab=cfg["vocab"])
    img_ids = tok.encode("".join(["<img>"] + ["<img_{:05d}>".format(i) for i in range(64)] + ["</img>"]))

    def make(max_len=256, **kw):
        m = llm_mod.LlamaForCausalLM(cfg, max_len=max_len, device="cpu", **kw)
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        return m

    ids_b = g["ids"] + [tok.encode("<img>")[0]]
    emb_b = torch.cat([g["embeds"], sd["model.embed_tokens.weight"][ids_b[-1]][None]])
    return dict(g=g, cfg=cfg, sd=sd, tok=tok, img_ids=img_ids, make=make, ids_b=ids_b, emb_b=emb_b)


def test_prefill_and_chunked_prefill_through_pages(host):
    g, m = host["g"], host["make"](kv_page_size=16)
    m.kv_alloc.shuffle(1)
    xs = m.prefill(g["embeds"])
    logits, hid = m.logits_all(xs)
    assert rel(logits, g["prefill_logits"]) < TOL and rel(hid, g["prefill_hidden"]) < TOL
    m.prefill(g["embeds"][:-8])
    tail, _ = m.logits_all(m.prefill(g["embeds"][-8:], pos0=g["embeds"].shape[0] - 8))      # keys / values gathered from scattered pages
    assert rel(tail, logits[-8:]) < TOL


@pytest.mark.parametrize("jump", [True, False])
def test_greedy_loop_matches_reference_golden(host, jump):
    g, m = host["g"], host["make"](kv_page_size=16)
    m.jump_forward = jump
    m.kv_alloc.shuffle(2)
    out = m.generate_greedy(g["ids"], g["embeds"], img_ids=host["img_ids"], max_new_tokens=16, use_graph=False)
    assert out.sequences[0][len(g["ids"]):].tolist() == g["text_gen_ids"] and rel(out.last_hidden_states, g["text_hidden"]) < TOL
    out = m.generate_greedy(host["ids_b"], host["emb_b"], img_ids=host["img_ids"], max_new_tokens=72, use_graph=False)
    assert out.sequences[0][len(host["ids_b"]):].tolist() == g["img_gen_ids"]
    assert out.n_generated == 72 and rel(out.last_hidden_states, g["img_hidden"]) < TOL
    assert sum(len(p) for p in m.slot_pages) + len(m.kv_alloc.free) == m.kv_alloc.n_pages


def test_jump_forward_bookkeeping(host):
    """how many token-loop steps remain, what lands in the sequence / harvest buffers, budget cut inside the span, surplus tokens dropped"""
    g, img_ids = host["g"], host["img_ids"]
    m = host["make"]()
    calls = {"n": 0}
    real = m._decode_step

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    m._decode_step = counting
    out = m.generate_greedy(host["ids_b"], host["emb_b"], img_ids=img_ids, max_new_tokens=72, use_graph=False)
    assert calls["n"] == 72 - 1 - 65                        # the 65 forced tokens rode in the prefill pass
    assert out.sequences[0][len(host["ids_b"]):].tolist() == g["img_gen_ids"]
    calls["n"] = 0
    out = m.generate_greedy(host["ids_b"], host["emb_b"], img_ids=img_ids, max_new_tokens=10, use_graph=False)     # budget ends inside the span
    assert calls["n"] == 0 and out.sequences[0][len(host["ids_b"]):].tolist() == img_ids[1:11] and out.last_hidden_states.shape[0] == 9
    assert rel(out.last_hidden_states, g["img_hidden"][:9]) < TOL
    out = m.generate_greedy(host["ids_b"], host["emb_b"], img_ids=img_ids, max_new_tokens=1, use_graph=False)
    assert out.sequences[0][len(host["ids_b"]):].tolist() == img_ids[1:2] and out.last_hidden_states.shape[0] == 0
    # lock-step batch where only one request jumps: the jumping slot runs surplus steps that are dropped again
    calls["n"] = 0
    outs = m.generate_greedy_batch([g["ids"], host["ids_b"]], [g["embeds"], host["emb_b"]], img_ids=img_ids, max_new_tokens=72, use_graph=False)
    assert calls["n"] == 71
    assert outs[0].sequences[0][len(g["ids"]):len(g["ids"]) + 16].tolist() == g["text_gen_ids"] and outs[0].n_generated == 72
    assert outs[1].sequences[0][len(host["ids_b"]):].tolist() == g["img_gen_ids"] and outs[1].n_generated == 72
    assert rel(outs[1].last_hidden_states, g["img_hidden"]) < TOL
    # without room for the surplus steps the shortcut is not taken (same result)
    tight = host["make"]()
    tight.max_len = len(host["ids_b"]) + 72 + 10
    tight._alloc_state(2)
    outs2 = tight.generate_greedy_batch([g["ids"], host["ids_b"]], [g["embeds"], host["emb_b"]], img_ids=img_ids, max_new_tokens=72, use_graph=False)
    assert [o.sequences.tolist() for o in outs2] == [o.sequences.tolist() for o in outs]


def test_lock_step_batch_of_ragged_prompts(host):
    g, img_ids = host["g"], host["img_ids"]
    m = host["make"](kv_page_size=16)
    outs = m.generate_greedy_batch([g["ids"], host["ids_b"], g["ids"][:20]], [g["embeds"], host["emb_b"], g["embeds"][:20]], img_ids=img_ids,
                                   max_new_tokens=72, use_graph=False)
    assert m.slots == 4
    assert outs[0].sequences[0][len(g["ids"]):len(g["ids"]) + 16].tolist() == g["text_gen_ids"]
    assert outs[1].sequences[0][len(host["ids_b"]):].tolist() == g["img_gen_ids"]
    single = m.generate_greedy(g["ids"][:20], g["embeds"][:20], img_ids=img_ids, max_new_tokens=72, use_graph=False)
    assert outs[2].sequences.tolist() == single.sequences.tolist()
    owned = [p for pages in m.slot_pages for p in pages]
    assert len(owned) == len(set(owned))


def test_eos_stops_a_sequence_and_hf_generate_surface(host):
    g, tok = host["g"], host["tok"]
    m = host["make"]()
    first = g["text_gen_ids"][0]
    seq = m.generate(input_ids=torch.tensor([g["ids"]]), inputs_embeds=g["embeds"][None], max_new_tokens=16, eos_token_id=g["text_gen_ids"][3], use_graph=False)
    assert seq[0].tolist() == g["ids"] + g["text_gen_ids"][:4]                   # stops at, and includes, the EOS (HF greedy_search)
    proc = llm_mod.AutoImageTokenGenerationProcessor(tok, num_img_gen_tokens=64)
    out = m.generate(input_ids=torch.tensor([host["ids_b"]]), inputs_embeds=host["emb_b"][None], output_hidden_states=True, return_dict_in_generate=True,
                     logits_processor=[proc], temperature=0.7, num_beams=1, max_new_tokens=72, top_p=0.5, do_sample=False, eos_token_id=None, use_graph=False)
    P = len(host["ids_b"])
    assert out.sequences[0][P:].tolist() == g["img_gen_ids"] and len(out.hidden_states) == 72
    last_hidden_states = torch.cat([h[-1] for h in out.hidden_states], dim=1)[0, P:, :]          # exactly seed_x.py:196-197
    assert rel(last_hidden_states, g["img_hidden"]) < TOL
    assert rel(out.hidden_states[0][-1][0, :P - 1], g["prefill_hidden"]) < TOL
    assert first == int(m.generate(input_ids=torch.tensor([g["ids"]]), inputs_embeds=g["embeds"][None], max_new_tokens=1, eos_token_id=None, use_graph=False)[0, -1])


def test_agent_generate_host_logic_vs_oracle(monkeypatch, host):
    """ContinuousLVLM.generate on the CPU double: input resampler + patch-position row, mask scatter, forced image span, hidden-state harvest,
    output resampler — ids / text exact and img_gen_feat <= 1e-3 against oracle/llm.py::lvlm_generate (multi-image prompt: 2 + 1 views)."""
    from oracle import llm as ollm
    from seedx_b200 import agent as agent_mod
    from seedx_b200 import demo
    from seedx_b200 import vit as vit_mod
    monkeypatch.setattr(agent_mod, "ops", fake_ops)
    monkeypatch.setattr(vit_mod, "ops", fake_ops)
    cfg, tok = host["cfg"], host["tok"]
    m = host["make"](max_len=512)
    vit_dim = 320
    agent_sd = synth.agent_state_dict(cfg["hidden"], vit_dim)
    agent = agent_mod.ContinuousLVLM.from_pretrained(llm=m, input_resampler=agent_mod.Resampler(8, cfg["hidden"], 2, vit_dim),
                                                     output_resampler=agent_mod.Resampler(8, vit_dim, 2, cfg["hidden"]), add_patch_pos=True, vit_down=True)
    agent.load_state_dict(agent_sd)
    N = 3
    image_embeds = synth.randn("agent_img_cpu", (N, 256, vit_dim))
    patch_pos = torch.tensor([[0.25, 0.5], [0.5, 0.5], [0.5, 0.5]])
    input_ids, ids_cmp_mask = demo.chat_prompt(tok, ["what changed?", "the sky", "draw it again"], views_per_image=[2, 1], force_image=True)
    assert int(ids_cmp_mask.sum()) == N * 64
    embeds_cmp_mask = torch.ones((N, 64), dtype=torch.bool)
    ref = ollm.lvlm_generate(host["sd"], agent_sd, cfg, tok, input_ids[0].tolist(), image_embeds, ids_cmp_mask[0], embeds_cmp_mask, patch_pos, 70)
    real = m.generate_greedy_batch
    monkeypatch.setattr(m, "generate_greedy_batch", lambda *a, **k: real(*a, **dict(k, use_graph=False)))
    out = agent.generate(tokenizer=tok, input_ids=input_ids, image_embeds=image_embeds, embeds_cmp_mask=embeds_cmp_mask, ids_cmp_mask=ids_cmp_mask,
                         patch_positions=patch_pos, max_new_tokens=70)
    assert out["ids"] == ref["ids"] and out["text"] == ref["text"]
    assert out["has_img_output"] and out["num_gen_imgs"] == 1 and tuple(out["img_gen_feat"].shape) == (1, 64, vit_dim)
    assert rel(out["img_gen_feat"], ref["img_gen_feat"]) < TOL


def test_mid_generation_span_jump_experimental(host):
    """SEEDX_JUMP_FORWARD_MID (off by default): a span the model opens by itself in the middle of its answer is finished by one chunked prefill.
    The tiny model is nudged to emit <img> (its lm_head row is aligned with one of its own hidden states); mode on == mode off."""
    g, tok, img_ids, sd = host["g"], host["tok"], host["img_ids"], dict(host["sd"])
    base = host["make"]()
    ref0 = base.generate_greedy(g["ids"], g["embeds"], img_ids=img_ids, max_new_tokens=24, use_graph=False)
    h = ref0.last_hidden_states[4]
    sd["lm_head.weight"] = sd["lm_head.weight"].clone()
    sd["lm_head.weight"][tok.tok2id["<img>"]] = (40.0 * h / h.pow(2).sum()).half().float()          # logit ~ 40 where the state is h

    def run(mid, ids_list, emb_list, n_new):
        m = llm_mod.LlamaForCausalLM(host["cfg"], max_len=256, device="cpu", kv_page_size=16)
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        m.jump_forward_mid = mid
        calls = {"n": 0}
        real = m._decode_step

        def counting(*a, **k):
            calls["n"] += 1
            return real(*a, **k)
        m._decode_step = counting
        return m.generate_greedy_batch(ids_list, emb_list, img_ids=img_ids, max_new_tokens=n_new, use_graph=False), calls["n"]

    off, n_off = run(False, [g["ids"]], [g["embeds"]], 100)
    on, n_on = run(True, [g["ids"]], [g["embeds"]], 100)
    gen = off[0].sequences[0][len(g["ids"]):].tolist()
    assert tok.tok2id["<img>"] in gen and tok.tok2id["</img>"] in gen, "the nudged model did not open an image span"
    assert on[0].sequences.tolist() == off[0].sequences.tolist() and on[0].n_generated == off[0].n_generated == 100
    assert rel(on[0].last_hidden_states, off[0].last_hidden_states) < TOL
    assert n_off == 99 and n_on <= 99 - 60                                  # at least one 64-token span left the token loop
    # lock-step pair: one request opens a span, the other (shorter prompt) does not necessarily; budget cut inside a span
    for n_new in (100, 12):
        off2, _ = run(False, [g["ids"], g["ids"][:20]], [g["embeds"], g["embeds"][:20]], n_new)
        on2, _ = run(True, [g["ids"], g["ids"][:20]], [g["embeds"], g["embeds"][:20]], n_new)
        for a, b in zip(on2, off2):
            assert a.sequences.tolist() == b.sequences.tolist() and a.n_generated == b.n_generated
            assert rel(a.last_hidden_states, b.last_hidden_states) < TOL


@pytest.mark.parametrize("page", [1, 16, 512])
def test_boundaries_exact_fit_page_sizes_and_eight_slots(host, page):
    """sequence that fills the cache to the last row, pages of one token / one page per sequence, 5 requests padded to 8 lock-step slots"""
    g, img_ids = host["g"], host["img_ids"]
    P = len(host["ids_b"])
    m = host["make"](max_len=P + 72, kv_page_size=page)                   # prompt + max_new_tokens == max_len exactly
    out = m.generate_greedy(host["ids_b"], host["emb_b"], img_ids=img_ids, max_new_tokens=72, use_graph=False)
    assert out.sequences[0][P:].tolist() == g["img_gen_ids"] and rel(out.last_hidden_states, g["img_hidden"]) < TOL
    from seedx_b200._lib import SeedxError
    with pytest.raises(SeedxError, match="exceeds the KV cache"):
        m.generate_greedy(host["ids_b"], host["emb_b"], img_ids=img_ids, max_new_tokens=73, use_graph=False)
    if page == 16:
        m8 = host["make"](max_len=160, kv_page_size=page)
        reqs = [g["ids"][:n] for n in (45, 7, 1, 33, 20)]
        embs = [g["embeds"][:n] for n in (45, 7, 1, 33, 20)]
        outs = m8.generate_greedy_batch(reqs, embs, img_ids=img_ids, max_new_tokens=16, use_graph=False)
        assert m8.slots == 8 and len(outs) == 5
        assert outs[0].sequences[0][45:].tolist() == g["text_gen_ids"]
        for r, e, o in zip(reqs, embs, outs):
            single = host["make"](max_len=160).generate_greedy(r, e, img_ids=img_ids, max_new_tokens=16, use_graph=False)
            assert o.sequences.tolist() == single.sequences.tolist()


def test_agent_generate_matches_the_reference_generate(monkeypatch, host):
    """product ContinuousLVLM.generate (CPU double) vs the reference's OWN ContinuousLVLM.generate output (tests/golden/agent_tiny.pt)"""
    from seedx_b200 import agent as agent_mod
    from seedx_b200 import vit as vit_mod
    monkeypatch.setattr(agent_mod, "ops", fake_ops)
    monkeypatch.setattr(vit_mod, "ops", fake_ops)
    g = torch.load(os.path.join(GOLD, "agent_tiny.pt"))
    cfg, tok = host["cfg"], host["tok"]
    m = host["make"](max_len=512)
    agent = agent_mod.ContinuousLVLM.from_pretrained(llm=m, input_resampler=agent_mod.Resampler(8, cfg["hidden"], 2, 320),
                                                     output_resampler=agent_mod.Resampler(8, 320, 2, cfg["hidden"]), add_patch_pos=True, vit_down=True)
    agent.load_state_dict(synth.agent_state_dict(cfg["hidden"], 320))
    real = m.generate_greedy_batch
    monkeypatch.setattr(m, "generate_greedy_batch", lambda *a, **k: real(*a, **dict(k, use_graph=False)))
    out = agent.generate(tokenizer=tok, input_ids=g["input_ids"], image_embeds=synth.randn("agent_golden_img", (3, 256, 320)),
                         embeds_cmp_mask=torch.ones((3, 64), dtype=torch.bool), ids_cmp_mask=g["ids_cmp_mask"], patch_positions=g["patch_pos"],
                         max_new_tokens=70, num_img_gen_tokens=64)
    assert out["text"] == g["text"] and out["has_img_output"] and out["num_gen_imgs"] == 1
    assert rel(out["img_gen_feat"], g["img_gen_feat"]) < TOL


def test_continuous_batching_equals_isolated_runs(host, monkeypatch):
    """seedx_b200.serving.ContinuousBatcher on the CPU double: five requests (two of them ending inside a forced image span, one hitting EOS early,
    ragged budgets) arrive staggered on two sequence slots; every request's ids and harvested hidden rows equal those of the same request run alone
    through generate_greedy, pages go back to the pool, and more requests than slots queue instead of failing."""
    from seedx_b200 import serving
    monkeypatch.setattr(serving, "ops", fake_ops)
    g, tok, sd = host["g"], host["tok"], host["sd"]
    emb_of = lambda ids: sd["model.embed_tokens.weight"][torch.tensor(ids)].float()  # noqa: E731
    base = list(g["ids"])
    reqs = [(base, g["embeds"], 12), (host["ids_b"], host["emb_b"], 70), (base[:9], g["embeds"][:9], 20), (base[:5] + [tok.encode("<img>")[0]], None, 30),
            (base[2:14], None, 7)]
    reqs = [(i, e if e is not None else emb_of(i), n) for i, e, n in reqs]
    solo = host["make"](max_len=256, kv_page_size=16)
    # an EOS that really occurs: the id request 2 generates third becomes the EOS id for everybody
    probe = solo.generate_greedy(reqs[2][0], reqs[2][1], img_ids=host["img_ids"], max_new_tokens=20, use_graph=False)
    eos = int(probe.sequences[0][len(reqs[2][0]) + 2])
    want = [solo.generate_greedy(i, e, img_ids=host["img_ids"], max_new_tokens=n, eos_id=eos, use_graph=False) for i, e, n in reqs]
    assert want[2].n_generated == 3                                            # stopped at (and including) the EOS
    m = host["make"](max_len=256, kv_page_size=16, kv_pages=40)
    cb = serving.ContinuousBatcher(m, slots=2, img_ids=host["img_ids"], eos_id=eos, use_graph=False)
    free0 = m.kv_alloc.n_free() if hasattr(m.kv_alloc, "n_free") else None
    got = cb.run(arrivals={0: [reqs[0]], 1: [reqs[1], reqs[2]], 4: [reqs[3]], 30: [reqs[4]]})
    assert sorted(got) == [0, 1, 2, 3, 4]
    for rid, w in enumerate(want):
        o = got[rid]
        assert o.n_generated == w.n_generated, rid
        assert torch.equal(o.sequences, w.sequences), rid
        assert o.last_hidden_states.shape == w.last_hidden_states.shape and (w.last_hidden_states.numel() == 0 or rel(o.last_hidden_states, w.last_hidden_states) < TOL), rid   # torch CPU matmuls of 1 vs 2 rows differ in summation order
    assert cb.idle() and all(r is None for r in cb.live)
    # lock-step would need max(steps) per wave of 2; continuous admission overlaps the long image span of request 1 with requests 2 and 3
    assert cb.steps < sum(w.n_generated for w in want)
    with pytest.raises(serving.SeedxError):
        cb.submit(base, g["embeds"], 10 ** 6)

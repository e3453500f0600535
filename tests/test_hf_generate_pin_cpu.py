"""Pins the greedy-loop semantics the reference relies on (seed_x.py:184-189: `self.llm.generate(input_ids=…, inputs_embeds=…, logits_processor=…,
output_hidden_states=True, return_dict_in_generate=True)`, HF GenerationMixin) against the transformers package INSTALLED here, and with it the oracle's
LLaMA forward against HF's own `LlamaForCausalLM`: the synthetic tiny checkpoint (HF key names, as the reference stores them) is loaded into
`transformers.LlamaForCausalLM`, `generate` is called exactly as the reference calls it — both `input_ids` and `inputs_embeds`, greedy, the reference's
logits-processor rule — and ids, stopping rule and the hidden rows the reference harvests must equal `oracle.llm.greedy_generate`.
The reference pins transformers 4.30.2 (requirements.txt); the installed 5.x cannot drive the reference's xformers model class, but its greedy loop
(processor call contract, argmax, EOS stop, what `sequences` / `hidden_states` hold when `inputs_embeds` is given) is the part restated in oracle/llm.py."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

transformers = pytest.importorskip("transformers")


def _hf_model(lc, lsd):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=lc["vocab"], hidden_size=lc["hidden"], intermediate_size=lc["ffn"], num_hidden_layers=lc["layers"],
                      num_attention_heads=lc["heads"], num_key_value_heads=lc["heads"], rms_norm_eps=lc["eps"], max_position_embeddings=512,
                      rope_theta=10000.0, tie_word_embeddings=False, attention_bias=False, bos_token_id=1, eos_token_id=2, pad_token_id=0)
    m = LlamaForCausalLM(cfg).float().eval()
    res = m.load_state_dict({k: v.float() for k, v in lsd.items()}, strict=False)
    assert not res.missing_keys and all(k.endswith("rotary_emb.inv_freq") for k in res.unexpected_keys), res
    return m


@pytest.mark.parametrize("prompt,n,eos_at", [("draw a cat<img>", 70, None), ("what is in the picture?", 24, None), ("what is in the picture?", 24, 5)])
def test_oracle_greedy_loop_equals_hf_generate(prompt, n, eos_at):
    from transformers import LogitsProcessorList
    from oracle import llm as ollm
    from seedx_b200 import synth
    from seedx_b200.llm import AutoImageTokenGenerationProcessor      # bit-equal to the reference class (test_oracle_golden.py)
    lc = synth.TINY_LLAMA
    lsd = synth.llama_state_dict(lc)
    tok = synth.SynthTokenizer(vocab=lc["vocab"])
    img_ids = tok.encode("".join(["<img>"] + ["<img_{:05d}>".format(i) for i in range(64)] + ["</img>"]))
    ids = [tok.bos_token_id] + tok.encode(prompt)
    emb = lsd["model.embed_tokens.weight"][torch.tensor(ids)].float()
    eos = 2
    if eos_at is not None:                                   # make the model "emit EOS": declare its own eos_at-th greedy token the EOS id
        free, _ = ollm.greedy_generate(lsd, lc, ids, emb, img_ids, n, eos_id=None)
        eos = free[eos_at]
        assert eos not in free[:eos_at]
    ref_ids, ref_hid = ollm.greedy_generate(lsd, lc, ids, emb, img_ids, n, eos_id=eos)
    m = _hf_model(lc, lsd)
    with torch.no_grad():
        out = m.generate(input_ids=torch.tensor([ids]), inputs_embeds=emb[None], max_new_tokens=n, do_sample=False,
                         logits_processor=LogitsProcessorList([AutoImageTokenGenerationProcessor(tok)]), eos_token_id=eos, pad_token_id=0,
                         return_dict_in_generate=True, output_hidden_states=True)
    got = out.sequences[0][len(ids):].tolist()              # the reference slices the prompt off the same way (seed_x.py:190)
    assert got == ref_ids
    if eos_at is not None:
        assert got[-1] == eos and len(got) == eos_at + 1
    # hidden rows: seed_x.py:196-197 concatenates the last layer's states of every step and drops the prompt part; the row that consumed generated
    # token j is the single row of step j + 1
    hf_rows = torch.stack([out.hidden_states[j + 1][-1][0, -1] for j in range(len(got) - 1)]) if len(got) > 1 else torch.zeros(0, lc["hidden"])
    assert hf_rows.shape == ref_hid.shape
    if len(got) > 1:
        assert ((hf_rows - ref_hid).norm() / ref_hid.norm()).item() < 1e-5

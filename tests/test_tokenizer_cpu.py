"""The tokenizer leg of the drop-in boundary with the REAL class: `configs/tokenizer/clm_llama_tokenizer_224loc_anyres.yaml` instantiates
`transformers.LlamaTokenizer.from_pretrained` (reference: eval_img2text_seed_x_i.py:76-77); tests elsewhere use the synthetic id map
(seedx_b200.synth.SynthTokenizer).  Here a genuine LlamaTokenizer directory is built offline (tests/real_tokenizer.py) and pushed through the YAML
factory, the prompt builders, the span mask and the agent's text / image-span split."""
import os

import pytest
import torch

import real_tokenizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tok(tmp_path_factory):
    from seedx_b200 import compat
    compat.install()
    import hydra
    from omegaconf import OmegaConf
    d = tmp_path_factory.mktemp("tok")
    real_tokenizer.build(str(d / "pretrained" / "cvlm_llama2_tokenizer_100img_and_224loc_addpatch"))
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "tokenizer", "clm_llama_tokenizer_224loc_anyres.yaml"))
    cwd = os.getcwd()
    os.chdir(d)                                   # the YAML names the directory relative to the project root, like the reference
    try:
        t = hydra.utils.instantiate(cfg)
    finally:
        os.chdir(cwd)
    return t


def test_real_llama_tokenizer_through_the_yaml_factory(tok):
    from transformers import LlamaTokenizer
    assert isinstance(tok, LlamaTokenizer) and len(tok) == 694 + 330 and tok.bos_token_id == 1 and tok.eos_token_id == 2
    ids = [tok.encode(t, add_special_tokens=False) for t in real_tokenizer.ADDED]
    assert all(len(i) == 1 for i in ids)                                     # every added token is ONE id ...
    flat = [i[0] for i in ids]
    assert flat == list(range(694, 694 + 330))                               # ... laid out behind the base vocabulary in the order they were added
    img = "".join("<img_{:05d}>".format(i) for i in range(64))
    s = "[INST] <patch>" + img + "</patch><img>" + img + "</img>Describe this image briefly. [/INST]\n"
    enc = tok.encode(s, add_special_tokens=False)
    assert tok.decode(enc, skip_special_tokens=False) == s                   # round trip incl. the trailing newline (byte fallback)


def test_prompt_builders_and_span_mask_with_the_real_tokenizer(tok):
    from seedx_b200 import demo
    ids, mask = demo.image_prompt(tok, 3, "Describe this image briefly.")
    first = tok.encode("<img_00000>", add_special_tokens=False)[0]
    assert ids.shape == mask.shape and int(mask.sum()) == 3 * 64
    assert ids[0][mask[0]].tolist() == list(range(first, first + 64)) * 3    # exactly the <img_k> rows, in order, once per view
    assert ids[0, 0].item() == tok.bos_token_id
    b_ids, b_mask = demo.image_prompt(tok, 0, "A cat on the road.", template=demo.BASE_GEN_PROMPT)
    assert b_ids[0, -1].item() == tok.encode("<img>", add_special_tokens=False)[0] and int(b_mask.sum()) == 0
    c_ids, c_mask = demo.chat_prompt(tok, ["what is in the image ?", "a cat .", "Make it red ."], views_per_image=[2, 1])
    assert int(c_mask.sum()) == 3 * 64


def test_agent_text_and_image_span_split_with_the_real_tokenizer(tok):
    """ContinuousLVLM._harvest (seed_x.py:191-223) on ids produced by the real tokenizer: the 64 rows before </img> are the image span, <img> and the
    span are stripped from the text, the logits-processor id list is the one the reference builds (generation.py:9-17)"""
    from seedx_b200.llm import AutoImageTokenGenerationProcessor
    proc = AutoImageTokenGenerationProcessor(tokenizer=tok, num_img_gen_tokens=64)
    img_str = "<img>" + "".join("<img_{:05d}>".format(i) for i in range(64)) + "</img>"
    assert proc.img_ids_list == tok.encode(img_str, add_special_tokens=False) and len(proc.img_ids_list) == 66
    gen = tok.encode("a red car " + img_str + " on the road", add_special_tokens=False)
    g = torch.tensor(gen)
    eoi, boi = tok.encode("</img>", add_special_tokens=False)[0], tok.encode("<img>", add_special_tokens=False)[0]
    e = int(torch.where(g == eoi)[0][0])
    assert g[e - 64:e].tolist() == proc.img_ids_list[1:65] and int(g[e - 65]) == boi
    keep = torch.ones_like(g, dtype=torch.bool)
    keep[e - 64:e] = False
    keep[g == boi] = False
    text = tok.decode(g[keep], skip_special_tokens=False)
    assert "a red car" in text and "on the road" in text and "<img_" not in text

"""Test helper: a REAL `transformers.LlamaTokenizer` directory built offline, in the layout of the reference's
`pretrained/cvlm_llama2_tokenizer_100img_and_224loc_addpatch` (configs/tokenizer/clm_llama_tokenizer_224loc_anyres.yaml): a SentencePiece BPE model
(trained here on a synthetic corpus, byte fallback like LLaMA's, <unk>/<s>/</s> = 0/1/2) plus the reference's added special tokens
<img> </img> <patch> </patch> <box_start> <box_end> <img_00000..00099> <loc-0..223>.  The LLaMA-2 vocabulary itself cannot be shipped; the class,
the file format, the added-token mechanics and the id layout (base vocabulary first, the 330 added tokens behind it) are the real ones."""
import os
import random

ADDED = ["<img>", "</img>", "<patch>", "</patch>", "<box_start>", "<box_end>"] + [f"<img_{i:05d}>" for i in range(100)] + [f"<loc-{i}>" for i in range(224)]
WORDS = ["the", "image", "shows", "a", "cat", "dog", "car", "red", "blue", "on", "road", "sky", "Describe", "this", "briefly", "Generate", "an", "edit", "Make", "it",
         "under", "sunset", "what", "is", "in", "Question", "Answer", ":", "[INST]", "[/INST]", "background", "there", "with", "connect", "advisor", "Sunday", "?", "."]


def build(path, base_vocab=694):
    """-> tokenizer directory at `path` with base_vocab + 330 ids (default 1024 = synth.TINY_LLAMA's vocabulary)"""
    import sentencepiece as spm
    from transformers import LlamaTokenizer
    os.makedirs(path, exist_ok=True)
    rnd = random.Random(0)
    corpus = os.path.join(path, "corpus.txt")
    with open(corpus, "w") as f:
        for _ in range(4000):
            f.write(" ".join(rnd.choice(WORDS) for _ in range(rnd.randint(3, 14))) + "\n")
        for _ in range(2000):      # enough character variety for the requested number of merges
            f.write(" ".join("".join(rnd.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rnd.randint(2, 9))) for _ in range(8)) + "\n")
    spm.SentencePieceTrainer.train(input=corpus, model_prefix=os.path.join(path, "tokenizer"), vocab_size=base_vocab, model_type="bpe", byte_fallback=True,
                                   character_coverage=1.0, unk_id=0, bos_id=1, eos_id=2, pad_id=-1, minloglevel=2)
    base_dir = os.path.join(path, "_base")
    os.makedirs(base_dir, exist_ok=True)
    os.replace(os.path.join(path, "tokenizer.model"), os.path.join(base_dir, "tokenizer.model"))
    tok = LlamaTokenizer.from_pretrained(base_dir)
    assert len(tok) == base_vocab
    assert tok.add_tokens(ADDED, special_tokens=True) == len(ADDED)
    tok.save_pretrained(path)
    return path

"""Where does the LLM stage of the bench go?  Times the pieces of ContinuousLVLM.generate_batch for 4 requests at full size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth, ops
from seedx_b200.llm import LlamaForCausalLM, LLAMA_13B
from seedx_b200.agent import ContinuousLVLM, Resampler

synth.set_device("cuda")
llm = LlamaForCausalLM(LLAMA_13B, max_len=1024); llm.load_state_dict(synth.llama_state_dict(LLAMA_13B))
agent = ContinuousLVLM.from_pretrained(llm=llm, input_resampler=Resampler(8, 5120, 32, 4096), output_resampler=Resampler(8, 4096, 32, 5120), add_patch_pos=True, vit_down=True)
agent.load_state_dict(synth.agent_state_dict(5120, 4096))
synth.set_device("cpu")
tok = synth.SynthTokenizer()
B, nv = 4, 2
img = "".join("<img_{:05d}>".format(i) for i in range(64))
s = ("<patch>" + img + "</patch>") * (nv - 1) + "<img>" + img + "</img>"
feats = torch.randn(B * nv, 256, 4096, device="cuda").half()
reqs = []
for b in range(B):
    ids = torch.tensor([tok.bos_token_id] + tok.encode("[INST] ") + tok.encode(s) + torch.randint(3, 30000, (32,)).tolist() + tok.encode(" [/INST]\n") + tok.encode("<img>"))
    first = tok.tok2id["<img_00000>"]
    mask = (ids >= first) & (ids < first + 64)
    reqs.append(dict(input_ids=ids.unsqueeze(0), image_embeds=feats[b * nv:(b + 1) * nv], embeds_cmp_mask=torch.ones((nv, 64), dtype=torch.bool),
                     ids_cmp_mask=mask.unsqueeze(0), patch_positions=torch.tensor([[0.5, 0.5]] * nv)))
def T():
    torch.cuda.synchronize(); return time.time()
for it in range(4):
    t0 = T()
    pairs = [agent._embed_request(r["input_ids"], r["image_embeds"], r["embeds_cmp_mask"], r["ids_cmp_mask"], r["patch_positions"]) for r in reqs]
    t1 = T()
    img_ids = tok.encode("<img>" + img + "</img>")
    outs = llm.generate_greedy_batch([p[0] for p in pairs], [p[1] for p in pairs], img_ids=img_ids, max_new_tokens=66, eos_id=2, suppress_eos=True)
    t2 = T()
    res = [agent._harvest(tok, o, p[0].numel(), 64) for o, p in zip(outs, pairs)]
    t3 = T()
    print(f"iter {it}: embed+input-resampler {1e3*(t1-t0):.1f} ms | generate_greedy_batch {1e3*(t2-t1):.1f} ms | harvest+output-resampler {1e3*(t3-t2):.1f} ms", flush=True)
# inside generate: prefill vs decode
P = pairs[0][0].numel()
t0 = T()
for s_ in range(4):
    xs = llm.prefill(pairs[s_][1].reshape(P, -1), slot=s_)
t1 = T()
print(f"4 prefills (P={P}): {1e3*(t1-t0):.1f} ms")

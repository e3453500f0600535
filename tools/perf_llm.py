"""Scratch perf probe: LLaMA-13B decode ms/token (CUDA-graph step) for 1 and 4 lock-step sequences + prefill time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth, ops
from seedx_b200.llm import LlamaForCausalLM, LLAMA_13B

synth.set_device("cuda")
m = LlamaForCausalLM(LLAMA_13B, max_len=1024)
m.load_state_dict(synth.llama_state_dict(LLAMA_13B))
synth.set_device("cpu")
P = 175
for B in (1, 4):
    ids = [torch.randint(3, 30000, (P,)).tolist() for _ in range(B)]
    embs = [m.get_input_embeddings()(torch.tensor(i)).view(P, -1) for i in ids]
    torch.cuda.synchronize()
    t0 = time.time()
    outs = m.generate_greedy_batch(ids, embs, max_new_tokens=66, suppress_eos=True, eos_id=2)
    torch.cuda.synchronize()
    t1 = time.time()
    # time the decode graph alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    hidden = torch.zeros((m.slots, 200, 5120), device="cuda")
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        m._decode_step(hidden, None, 2, True)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    e0.record(); xs = m.prefill(embs[0], slot=0); e1.record(); torch.cuda.synchronize()
    print(f"B={B}: generate(66 tok) wall {1e3*(t1-t0):.0f} ms; decode step {ms:.2f} ms -> weights {26.04/ms*1e3/1e3:.2f} TB/s "
          f"({26.04/ms/6.4852*100:.0f}% of HBM peak), {B/ms*1e3:.0f} tok/s; prefill P={P}: {e0.elapsed_time(e1):.1f} ms", flush=True)

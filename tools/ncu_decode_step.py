"""LLaMA-13B (random init) greedy decoding run eagerly (no CUDA graph) inside a cudaProfilerStart/Stop range: prefill of a 182-row prompt + STEPS token
steps of B lock-step sequences, for the launch list of the HBM-bound token loop (gemv_mma_kernel, decode_attn_kernel, logits_argmax_kernel ...):
   ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
       --cache-control none --csv --log-file X python tools/ncu_decode_step.py
env: B (sequences, default 1), STEPS (default 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seedx_b200 import synth
from seedx_b200.llm import LLAMA_13B, LlamaForCausalLM
B, STEPS = int(os.environ.get("B", "1")), int(os.environ.get("STEPS", "4"))
synth.set_device("cuda"); sd = synth.llama_state_dict(LLAMA_13B); synth.set_device("cpu")
m = LlamaForCausalLM(LLAMA_13B, max_len=512); m.load_state_dict(sd); del sd
g = torch.Generator().manual_seed(5)
ids = [torch.randint(3, 30000, (182,), generator=g) for _ in range(B)]
emb = [m.get_input_embeddings()(i)[0] for i in ids]
run = lambda: m.generate_greedy_batch(ids, emb, img_ids=None, max_new_tokens=STEPS + 1, eos_id=None, suppress_eos=True, use_graph=False)
run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", STEPS, "decode steps")

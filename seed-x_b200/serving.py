"""Continuous batching of the LLaMA token loop: requests are admitted into free sequence slots BETWEEN replays of the
lock-step decode graph instead of only at the start of a batch.

The reference decodes one request at a time with HF `generate` (src/models/mllm/seed_x.py:184-189); `LlamaForCausalLM.generate_greedy_batch`
already streams the weights once per step for up to 8 requests that start together.  What makes admission in the middle cheap here is that every
piece of per-request state the captured graph reads is a row of a static device buffer (`seq`, `state`, the page table, the harvest buffer): a
new request is a prefill into one slot's pages plus a few row writes, the graph is neither re-captured nor aware of it.  A retired or idle slot is
parked by a non-zero `state[slot][1]` (seedx_logits_argmax): it keeps riding along in the batched GEMVs — the weights are read anyway — but its
sequence, KV length and harvest row no longer move.

Results are those of running every request alone in the same engine: slots never read each other's rows
(`tests/test_llm_host_cpu.py::test_continuous_batching_equals_isolated_runs`, `tests/test_llm_gpu.py::test_continuous_batching_gpu`)."""
import collections

import torch

from . import _lib, ops
from ._lib import SeedxError
from .llm import GreedyOutput


class ContinuousBatcher:
    PARKED = -1          # state[slot][1] of a slot without a live request

    def __init__(self, llm, slots=4, img_ids=None, eos_id=None, use_graph=True):
        if slots not in (1, 2, 4, 8):
            raise SeedxError("slots must be 1, 2, 4 or 8 (the batched GEMV variants)")
        self.llm, self.n_slots, self.eos_id, self.use_graph = llm, slots, eos_id, use_graph
        dev = llm.device
        llm._alloc_state(slots)
        self.img_key = tuple(int(i) for i in img_ids) if img_ids is not None else None
        self.img_dev = torch.tensor(list(self.img_key), dtype=torch.int32, device=dev) if self.img_key is not None else None
        llm._img_key, llm._img_dev = self.img_key, self.img_dev
        llm._hidden = torch.zeros((slots, llm.max_len, llm.cfg["hidden"]), device=dev, dtype=torch.float32)
        self.hidden = llm._hidden
        self.pending = collections.deque()
        self.live = [None] * slots            # per slot: dict(rid, P, ahead, budget)
        self.graph = None
        self.next_rid = 0
        self.steps = 0
        st = torch.zeros((slots, 4), dtype=torch.int32)
        st[:, 0], st[:, 1], st[:, 3] = 1, self.PARKED, 1          # parked: length 1 (a valid cache row), nothing appended
        llm.state.copy_(st)
        for s in range(slots):
            llm.reserve_kv(s, 1, fresh=True)

    # ---- requests --------------------------------------------------------------------------------------------------------------
    def submit(self, input_ids, inputs_embeds, max_new_tokens):
        """queue a request (prompt ids [P], prompt embeddings [P, D]); returns its id"""
        ids = torch.as_tensor(input_ids).reshape(-1)
        if ids.numel() + max_new_tokens > self.llm.max_len:
            raise SeedxError("prompt + max_new_tokens exceeds the KV cache")
        rid = self.next_rid
        self.next_rid += 1
        self.pending.append(dict(rid=rid, ids=ids, emb=inputs_embeds.reshape(ids.numel(), -1), budget=int(max_new_tokens)))
        return rid

    def _admit(self, slot, req):
        """prefill one slot and write its rows of the static state: exactly what generate_greedy_batch does for every slot at the start of a batch,
        including the jump-forward over a forced image span"""
        llm, dev = self.llm, self.llm.device
        ids, emb = req["ids"], req["emb"].to(dev, torch.float32)
        P = ids.numel()
        ahead = []
        last = int(ids[-1])
        if llm.jump_forward and self.img_key is not None and last in self.img_key[:-1]:
            ahead = list(self.img_key[self.img_key.index(last) + 1:])[: max(req["budget"] - 1, 0)]
        llm.seq[slot, :P].copy_(ids.to(dev, torch.int32))
        if ahead:
            a_t = torch.tensor(ahead, dtype=torch.int32, device=dev)
            llm.seq[slot, P:P + len(ahead)].copy_(a_t)
            emb = torch.cat([emb, llm.get_input_embeddings()(a_t)[0]], dim=0)
        xs = llm.prefill(emb, slot=slot)                              # fresh=True inside: the slot's old pages go back to the pool first
        llm.reserve_kv(slot, P + req["budget"])
        na = len(ahead)
        ops.gemv(llm.lm_head, xs[P + na - 1], llm.logits[slot], rms_w=llm.norm, eps=llm.cfg["eps"])
        if na:
            ops.layernorm(xs[P:P + na], llm.norm, None, llm.cfg["eps"], out=self.hidden[slot, :na], rms=True)
        llm.state[slot].copy_(torch.tensor([P + na, 0, na, P], dtype=torch.int32))
        # first generated token of this slot only (the other slots' rows must not move): the batched kernel on a one-row view
        ops.logits_argmax(llm.logits[slot:slot + 1], self.img_dev, llm.seq[slot:slot + 1], llm.state[slot:slot + 1], self.eos_id, False)
        self.live[slot] = dict(rid=req["rid"], P=P, budget=req["budget"])

    def _retire(self, slot, n_gen):
        llm, r = self.llm, self.live[slot]
        P = r["P"]
        seq = llm.seq[slot, :P + n_gen].to(torch.int64).cpu().unsqueeze(0)
        out = GreedyOutput(seq, self.hidden[slot, :max(n_gen - 1, 0)].clone(), n_gen)
        out.prefill_hidden = None
        llm.state[slot].copy_(torch.tensor([1, self.PARKED, 0, 1], dtype=torch.int32))       # park: the next admission overwrites everything
        self.live[slot] = None
        return r["rid"], out

    # ---- the loop --------------------------------------------------------------------------------------------------------------
    def _replay(self):
        llm = self.llm
        if not self.use_graph:
            llm._decode_step(self.hidden, self.img_dev, self.eos_id, False)
            return
        if self.graph is None:
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            # capturing executes nothing: the state the first replay reads is the one the admissions below have written
            with torch.cuda.graph(self.graph, stream=s_):
                llm._decode_step(self.hidden, self.img_dev, self.eos_id, False)
            self.graph.n_kernels = _lib.launch_count() - n0
            torch.cuda.current_stream().wait_stream(s_)
        self.graph.replay()
        _lib.note_replay(self.graph.n_kernels)

    def step(self):
        """admit waiting requests into free slots, retire what is complete, run one token step for all slots.  Returns the finished
        (request id, GreedyOutput) pairs."""
        llm = self.llm
        for s in range(self.n_slots):
            if self.live[s] is None and self.pending:
                self._admit(s, self.pending.popleft())
        done = []
        st = llm.state.cpu().tolist()                   # one small read per step: lengths and EOS marks of all slots
        for s in range(self.n_slots):
            r = self.live[s]
            if r is None:
                continue
            n_gen = st[s][2]
            if st[s][1] > 0:                            # EOS produced: stop at (and include) it, like HF greedy_search
                done.append(self._retire(s, min(st[s][1], r["budget"])))
            elif n_gen >= r["budget"]:
                done.append(self._retire(s, r["budget"]))
        if any(r is not None for r in self.live):
            self._replay()
            self.steps += 1
        return done

    def idle(self):
        return not self.pending and all(r is None for r in self.live)

    def run(self, arrivals=None):
        """drive the loop until everything submitted has finished.  arrivals: optional {step index: [(ids, emb, max_new_tokens), ...]} submitted when
        the loop reaches that step (staggered arrivals for tests / load generators).  Returns {request id: GreedyOutput}."""
        arrivals = dict(arrivals or {})
        results, i = {}, 0
        while arrivals or not self.idle():
            for a in arrivals.pop(i, []):
                self.submit(*a)
            for rid, out in self.step():
                results[rid] = out
            i += 1
            if i > 100000:
                raise SeedxError("ContinuousBatcher.run: no progress")
        return results

"""Optional fine-grained stage timing: ``mark(name)`` records a CUDA event on the current stream when tracing is on
(bench.py turns it on for one extra untimed step; it is off inside every timed region, where it would only add event records)."""
import torch

_on = False
_marks = []


def enable(flag=True):
    global _on
    _on = bool(flag)
    _marks.clear()


def mark(name):
    if _on:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        _marks.append((name, e))


def summary():
    """ms between consecutive marks, summed per name of the *later* mark (= the section that ended there)."""
    torch.cuda.synchronize()
    out = {}
    for (_, a), (n, b) in zip(_marks, _marks[1:]):
        out[n] = out.get(n, 0.0) + a.elapsed_time(b)
    return out

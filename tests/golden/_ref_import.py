"""Import the UNMODIFIED reference modules from /root/reference on CPU (only in the build container; the GPU box has no
/root/reference).  Three import stubs are needed: deepspeed, transformers.deepspeed, xformers.ops."""
import importlib
import sys
import types

import torch
import torch.nn.functional as F

REF = "/root/reference"


def install_stubs():
    if "deepspeed" not in sys.modules:
        ds = types.ModuleType("deepspeed")
        ds.zero = types.SimpleNamespace(GatheredParameters=None)
        sys.modules["deepspeed"] = ds
    if "transformers.deepspeed" not in sys.modules:
        td = types.ModuleType("transformers.deepspeed")
        td.is_deepspeed_zero3_enabled = lambda: False
        sys.modules["transformers.deepspeed"] = td
    if "xformers" not in sys.modules:
        xf = types.ModuleType("xformers")
        xo = types.ModuleType("xformers.ops")

        class LowerTriangularMask:  # marker type, as in xformers
            pass

        def memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None):
            # q,k,v: [B,S,H,D] -> exact softmax attention (xformers semantics)
            causal = isinstance(attn_bias, LowerTriangularMask)
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal, scale=scale)
            return o.transpose(1, 2)

        xo.memory_efficient_attention = memory_efficient_attention
        xo.LowerTriangularMask = LowerTriangularMask
        xf.ops = xo
        sys.modules["xformers"] = xf
        sys.modules["xformers.ops"] = xo
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # the reference's `src` is a namespace package (no __init__.py) while this repo ships a regular `src` shim package, which would win
    # the import regardless of sys.path order: bind `src` to the reference tree explicitly
    cur = sys.modules.get("src")
    if cur is None or list(getattr(cur, "__path__", [])) != [REF + "/src"]:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        pkg = types.ModuleType("src")
        pkg.__path__ = [REF + "/src"]
        sys.modules["src"] = pkg


def ref_module(dotted):
    install_stubs()
    m = importlib.import_module(dotted)
    assert m.__file__.startswith(REF + "/"), f"{dotted} resolved to {m.__file__}, not the reference"
    return m

"""Minimal stand-ins for the third-party helpers the reference entry scripts import but this image does not ship
(hydra, omegaconf, pyrootutils, diffusers, peft; SURVEY.md §5 'Config / flag system', §8b).  ``install()`` registers them in
sys.modules only when the real package is missing, so `import hydra` / `from omegaconf import OmegaConf` /
`from diffusers import AutoencoderKL, UNet2DConditionModel, EulerDiscreteScheduler[, Transformer2DModel]` in src/inference/eval_*.py resolve.
"""
import importlib
import os
import sys
import types

import yaml


class DictConfig(dict):
    """attribute-accessible dict (the subset of omegaconf.DictConfig the scripts use)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _wrap(o):
    if isinstance(o, dict):
        return DictConfig({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_wrap(v) for v in o]
    return o


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def create(obj):
        return _wrap(obj)


def _locate(dotted):
    parts = dotted.split(".")
    for i in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:i]))
        except ImportError:
            continue
        for p in parts[i:]:
            obj = getattr(obj, p)
        return obj
    raise ImportError(f"cannot locate {dotted!r}")


def instantiate(cfg, *args, **overrides):
    """hydra.utils.instantiate for `_target_` configs: nested `_target_` dicts are instantiated first (recursive mode),
    call-time keyword overrides win over YAML values (eval_img2text_seed_x_i.py:88,92,104)."""
    if not isinstance(cfg, dict) or "_target_" not in cfg:
        raise ValueError("instantiate() needs a mapping with a _target_ key")
    kwargs = {}
    for k, v in cfg.items():
        if k in ("_target_", "_recursive_", "_convert_", "_partial_"):
            continue
        kwargs[k] = instantiate(v) if isinstance(v, dict) and "_target_" in v else v
    kwargs.update(overrides)
    return _locate(cfg["_target_"])(*args, **kwargs)


def setup_root(search_from, indicator=".project-root", pythonpath=True, **kw):
    """pyrootutils.setup_root: walk up to the directory holding the marker file, put it on sys.path, chdir-free."""
    d = os.path.abspath(search_from if os.path.isdir(search_from) else os.path.dirname(search_from))
    while True:
        if os.path.exists(os.path.join(d, indicator)):
            if pythonpath and d not in sys.path:
                sys.path.insert(0, d)
            os.environ.setdefault("PROJECT_ROOT", d)
            return d
        nd = os.path.dirname(d)
        if nd == d:
            raise FileNotFoundError(f"{indicator} not found above {search_from}")
        d = nd


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except ImportError:
        return True


def install():
    if _missing("omegaconf"):
        m = types.ModuleType("omegaconf")
        m.OmegaConf, m.DictConfig = OmegaConf, DictConfig
        sys.modules["omegaconf"] = m
    if _missing("hydra"):
        h, u = types.ModuleType("hydra"), types.ModuleType("hydra.utils")
        u.instantiate = instantiate
        h.utils = u
        sys.modules["hydra"], sys.modules["hydra.utils"] = h, u
    if _missing("pyrootutils"):
        p = types.ModuleType("pyrootutils")
        p.setup_root = setup_root
        sys.modules["pyrootutils"] = p
    if _missing("peft"):
        from . import lora
        pm = types.ModuleType("peft")
        pm.LoraConfig = lora.LoraConfig
        sys.modules["peft"] = pm
    if _missing("diffusers"):
        from . import sdxl
        d = types.ModuleType("diffusers")
        d.AutoencoderKL, d.UNet2DConditionModel, d.EulerDiscreteScheduler = sdxl.AutoencoderKL, sdxl.UNet2DConditionModel, sdxl.EulerDiscreteScheduler
        # eval_img2edit_seed_x_edit.py:8 imports the name without ever instantiating it
        d.Transformer2DModel = sdxl.Transformer2DModel
        sys.modules["diffusers"] = d

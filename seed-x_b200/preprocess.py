"""Host-side image pre-processing (PIL / numpy; stays on the CPU exactly like the reference's):

  * get_transform          <- src/processer/transforms.py:5-20 ('clip': Resize((S,S)) bilinear -> [0,1] CHW -> CLIP mean/std)
  * process_anyres_image   <- src/inference/any_res.py:158-201 (grid choice, 448-px tiles + one global view, patch positions)
"""
import ast

import numpy as np
import torch
from PIL import Image

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_STATS = {"clip": (CLIP_MEAN, CLIP_STD), "clipa": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))}


class _ClipTransform:
    def __init__(self, size, keep_ratio, mean, std):
        self.size, self.keep_ratio = size, keep_ratio
        self.mean = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
        self.std = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)

    def __call__(self, img):
        img = img.convert("RGB")
        S = self.size
        if self.keep_ratio:                       # Resize(S) on the short side + CenterCrop(S)
            w, h = img.size
            if w <= h:
                nw, nh = S, int(S * h / w)
            else:
                nw, nh = int(S * w / h), S
            img = img.resize((nw, nh), Image.BILINEAR)
            left, top = int(round((nw - S) / 2.0)), int(round((nh - S) / 2.0))
            img = img.crop((left, top, left + S, top + S))
        else:
            img = img.resize((S, S), Image.BILINEAR)
        a = np.asarray(img, dtype=np.uint8).astype(np.float32).transpose(2, 0, 1) / 255.0
        return torch.from_numpy((a - self.mean) / self.std)


def get_transform(type="clip", keep_ratio=True, image_size=224):
    if type not in _STATS:
        raise ValueError(f"transform type {type!r} is not used by the inference configs (configs/processer/qwen_448_transform.yaml)")
    mean, std = _STATS[type]
    return _ClipTransform(image_size, keep_ratio, mean, std)


def _fit_by_coverage(size, grids):
    """largest effective (non-upscaled) resolution, ties -> least wasted area (any_res.py:9-36)"""
    ow, oh = size
    best, best_eff, best_waste = None, 0, float("inf")
    for w, h in grids:
        s = min(w / ow, h / oh)
        eff = min(int(ow * s) * int(oh * s), ow * oh)
        waste = w * h - eff
        if eff > best_eff or (eff == best_eff and waste < best_waste):
            best, best_eff, best_waste = (w, h), eff, waste
    return best


def _fit_by_aspect(size, grids):
    """closest aspect ratio, ties -> closest area (any_res.py:39-68)"""
    ow, oh = size
    ar0, area0 = oh / ow, ow * oh
    best, key = None, (float("inf"), float("inf"))
    for w, h in grids:
        ar, area = h / w, w * h
        k = (max(ar, ar0) / min(ar, ar0), max(area, area0) / min(area, area0))
        if k[0] < key[0] or (k[0] == key[0] and k[1] < key[1]):
            best, key = (w, h), k
    return best


def process_anyres_image(image, image_transform, grid_pinpoints, base_image_size):
    """-> (views [N,3,S,S] = row-major tiles of the resized image + one global view, patch_pos [N,2] = tile centres in [0,1])."""
    grids = grid_pinpoints if isinstance(grid_pinpoints, list) else ast.literal_eval(grid_pinpoints)
    a, b = _fit_by_coverage(image.size, grids), _fit_by_aspect(image.size, grids)
    W, H = b if a[0] * a[1] > b[0] * b[1] else a
    S = base_image_size
    canvas = image.resize((W, H))
    views = [canvas.crop((x, y, x + S, y + S)) for y in range(0, H, S) for x in range(0, W, S)]
    views.append(image.resize((S, S)))
    gx, gy = W // S, H // S
    pos = [[(ix + 0.5) / gx, (iy + 0.5) / gy] for iy in range(gy) for ix in range(gx)] + [[0.5, 0.5]]
    return torch.stack([image_transform(v) for v in views], dim=0), torch.tensor(pos, dtype=torch.float32)

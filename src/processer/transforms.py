"""`src.processer.transforms.get_transform` (reference: src/processer/transforms.py:5-83)."""
from seedx_b200.preprocess import get_transform  # noqa: F401

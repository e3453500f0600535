"""`from any_res import process_anyres_image` (reference: src/inference/any_res.py:158-201)."""
from seedx_b200.preprocess import process_anyres_image  # noqa: F401

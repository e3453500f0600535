"""`src.models.tokenizer.discrete_models.DiscreteModleIdentity` (reference: src/models/tokenizer/discrete_models.py:7-17)."""
from seedx_b200.adapter import DiscreteModleIdentity  # noqa: F401

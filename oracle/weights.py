"""Synthetic checkpoint generator + model constants, shared with the benchmark: re-exported from
advancedliteratemachinery_b200/synthetic.py (pure data generation, no model arithmetic)."""
from advancedliteratemachinery_b200.synthetic import *  # noqa: F401,F403
from advancedliteratemachinery_b200.synthetic import _Gen, _linear, _ln  # noqa: F401

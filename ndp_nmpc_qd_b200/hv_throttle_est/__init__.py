"""Drop-in for ndp_nmpc/scripts/hv_throttle_est (reference import: nmpc_node.py:32)."""
from .hover_throttle_estimator import AlphaFilter, Differentiator, HoverThrottleEstimator  # noqa: F401

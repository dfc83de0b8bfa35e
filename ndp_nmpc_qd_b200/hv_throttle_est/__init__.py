"""Drop-in for ndp_nmpc/scripts/hv_throttle_est (reference import: nmpc_node.py:32)."""
if not __package__ or "." not in __package__:
    # imported by the reference's bare name (PYTHONPATH=<repo>/ndp_nmpc_qd_b200): become the real module
    import os as _os
    import sys as _sys

    _root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    if _root not in _sys.path:
        _sys.path.append(_root)
    from ndp_nmpc_qd_b200.dropin import alias as _alias

    _alias(__name__)
else:
    from .hover_throttle_estimator import AlphaFilter, Differentiator, HoverThrottleEstimator  # noqa: F401

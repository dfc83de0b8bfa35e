"""Drop-in for dop_sim/scripts/quadrotor (reference import: dop_qd_node.py:22 `from quadrotor import MulQuadrotors`):
the batched CUDA plant of ndp_nmpc_qd_b200.dop_sim under the simulator node's import name."""
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
    from ..dop_sim.mul_quadrotors import MulQuadrotors  # noqa: F401

"""Drop-in for ndp_nmpc/scripts/dnwash_nn_est (reference import: nmpc_node.py:31)."""
from .downwash_nn import DownwashNN  # noqa: F401

"""Drop-in for ndp_nmpc/scripts/nmpc_ctl (reference import: nmpc_node.py:29)."""
from .nmpc_body_rate_ctl import NMPCBodyRateController  # noqa: F401

"""Drop-in for ndp_nmpc/scripts/ndp_nmpc_ctl (reference import: nmpc_node.py:30)."""
from .ndp_nmpc_body_rate_ctl import NDPNMPCBodyRateController  # noqa: F401

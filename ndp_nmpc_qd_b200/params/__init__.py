"""Constants of the controller, mirroring ndp_nmpc/scripts/params/ of the reference."""
from . import downwash_params, estimator_params, fhnp_params, nmpc_params  # noqa: F401

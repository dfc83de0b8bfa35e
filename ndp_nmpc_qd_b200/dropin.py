"""Makes the reference's own import statements resolve to this package.

The reference's nodes import their controller and estimators by bare package name
(ndp_nmpc/scripts/nmpc_node.py:29-34, ndp_nmpc_leader_node.py:20-25, nmpc_follower_node.py:23,
dop_sim/scripts/dop_qd_node.py:22):

    from nmpc_ctl import NMPCBodyRateController
    from ndp_nmpc_ctl import NDPNMPCBodyRateController
    from hv_throttle_est import HoverThrottleEstimator        (follower: AlphaFilter)
    from params import nmpc_params as CP, estimator_params as EP      (leader: downwash_params as DP)
    from dnwash_nn_est import DownwashNN
    from quadrotor import MulQuadrotors

Two ways to get there:
  * PYTHONPATH=<repo>/ndp_nmpc_qd_b200 -- the sub-packages are importable under their bare names; each one
    notices that it is being imported without its parent and aliases itself to the real
    ndp_nmpc_qd_b200.<name> module (one module object per package, so isinstance dispatch keeps working).
  * install() / `python -m ndp_nmpc_qd_b200.dropin <node.py> [args]` -- registers the aliases in sys.modules
    BEFORE the node runs.  This is the one that works for the unmodified ROS nodes: they prepend their own
    scripts directory to sys.path (nmpc_node.py:13-14), which would shadow any PYTHONPATH entry, but an
    entry already in sys.modules wins over the path search.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys

ALIASES = ("params", "nmpc_ctl", "ndp_nmpc_ctl", "dnwash_nn_est", "hv_throttle_est", "quadrotor")
_PKG = "ndp_nmpc_qd_b200"


def alias(bare_name: str):
    """Called from a sub-package's __init__ when it was imported by its bare name: returns the real module and
    puts it (and its sub-modules) in sys.modules under the bare name."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.append(root)
    real = importlib.import_module(f"{_PKG}.{bare_name}")
    sys.modules[bare_name] = real
    prefix = f"{_PKG}.{bare_name}."
    for name, mod in list(sys.modules.items()):
        if name.startswith(prefix) and mod is not None:
            sys.modules[bare_name + "." + name[len(prefix):]] = mod
    return real


def install(names=ALIASES) -> None:
    """Register every drop-in package under the reference's bare import name."""
    for n in names:
        alias(n)


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m ndp_nmpc_qd_b200.dropin <node.py> [node args...]")
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")


if __name__ == "__main__":
    main()

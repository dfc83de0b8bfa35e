"""The reference's own import statements must resolve to this package.

nmpc_node.py:29-34, ndp_nmpc_leader_node.py:20-25, nmpc_follower_node.py:23 and dop_qd_node.py:22 import the
controller / estimator / plant packages by bare name; these tests run exactly those lines in a clean
interpreter (i) with PYTHONPATH=<repo>/ndp_nmpc_qd_b200 and (ii) through the launcher
`python -m ndp_nmpc_qd_b200.dropin node.py`, where the node -- like the reference's -- prepends its own
directory (which holds decoy packages of the same names) to sys.path."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ndp_nmpc_qd_b200")

# verbatim from the reference nodes
NODE_IMPORTS = textwrap.dedent("""
    from nmpc_ctl import NMPCBodyRateController
    from ndp_nmpc_ctl import NDPNMPCBodyRateController
    from hv_throttle_est import HoverThrottleEstimator

    from params import nmpc_params as CP, estimator_params as EP
    from params import downwash_params as DP, nmpc_params as CP
    from dnwash_nn_est import DownwashNN
    from hv_throttle_est import AlphaFilter
    from quadrotor import MulQuadrotors
""")

CHECKS = textwrap.dedent("""
    import sys
    import ndp_nmpc_qd_b200.nmpc_ctl as real_ctl, ndp_nmpc_qd_b200.ndp_nmpc_ctl as real_ndp
    import ndp_nmpc_qd_b200.dnwash_nn_est as real_nn, ndp_nmpc_qd_b200.dop_sim as real_sim
    assert NMPCBodyRateController is real_ctl.NMPCBodyRateController
    assert NDPNMPCBodyRateController is real_ndp.NDPNMPCBodyRateController
    assert DownwashNN is real_nn.DownwashNN and MulQuadrotors is real_sim.MulQuadrotors
    # nmpc_node.py:203-208 dispatches with isinstance, plain controller first: the classes must be siblings
    assert not issubclass(NDPNMPCBodyRateController, NMPCBodyRateController)
    assert CP.N_node == 20 and CP.ts_nmpc == 0.02 and CP.long_list_size == 101 and EP.k_throttle_init == 50 and DP.r_horiz == 1.0
    assert HoverThrottleEstimator(EP.ts_est).update(0.0, 0.27434003169930943)[0] > 50.0
    print("DROPIN-OK")
""")


def _run(args, env_extra, cwd):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    env.update(env_extra)
    return subprocess.run([sys.executable] + args, capture_output=True, text=True, cwd=cwd, env=env, timeout=300)


def test_reference_import_lines_with_pythonpath(tmp_path):
    script = tmp_path / "imports.py"
    script.write_text(NODE_IMPORTS + CHECKS)
    r = _run([str(script)], {"PYTHONPATH": PKG}, str(tmp_path))
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout + r.stderr


def test_launcher_beats_the_nodes_own_sys_path(tmp_path):
    """The reference's nodes do `sys.path.insert(0, current_path)` before importing (nmpc_node.py:13-14), so their own
    scripts directory would win over PYTHONPATH; the launcher's sys.modules aliases must win over that."""
    for name in ("nmpc_ctl", "ndp_nmpc_ctl", "dnwash_nn_est", "hv_throttle_est", "params", "quadrotor"):
        d = tmp_path / name
        d.mkdir()
        (d / "__init__.py").write_text("raise ImportError('decoy: the reference package was imported, not the drop-in')\n")
    node = tmp_path / "fake_node.py"
    node.write_text(textwrap.dedent("""
        import sys
        import os

        current_path = os.path.abspath(os.path.dirname(__file__))
        sys.path.insert(0, current_path)
    """) + NODE_IMPORTS + CHECKS + "assert sys.argv[1:] == ['--flag'], sys.argv\n")
    r = _run(["-m", "ndp_nmpc_qd_b200.dropin", str(node), "--flag"], {"PYTHONPATH": ROOT}, str(tmp_path))
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout + r.stderr

"""ORACLE package -- CPU restatements of the reference's hot path.

Test infrastructure only: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(ndp_nmpc_qd_b200) never imports it.
"""

"""ORACLE package -- CPU restatement of the reference hot path (test infrastructure only).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs ONLY.  See oracle/lbm_oracle.py for the parity-pin statement.
"""

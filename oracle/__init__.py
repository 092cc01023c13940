"""ORACLE package — test infrastructure.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import it."""

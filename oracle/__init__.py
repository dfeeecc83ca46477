"""TEST INFRASTRUCTURE -- the parity checker, never the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package (see cg_oracle.h).
"""

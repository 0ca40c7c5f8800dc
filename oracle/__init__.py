"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/bgw_oracle.h).

Nothing under abmarl_b200/ imports this package; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg do, and only as the checker / reported baseline.
"""

"""CPU parity oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this package.  The product (plan_b200) never does.
"""

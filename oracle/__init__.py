"""TEST INFRASTRUCTURE ONLY: CPU oracle for the batched condensed-MPC path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this package.  The product package never does.
"""

"""CPU oracle for the daisyRec BPR hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (daisyrec_b200/) never does.
"""

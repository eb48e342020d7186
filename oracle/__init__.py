"""CPU oracle for the theanet hot path -- TEST INFRASTRUCTURE, never imported by theanet_b200/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
"""

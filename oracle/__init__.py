"""CPU oracle of the clip-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import or execute anything in this package; the product (clip-retrieval_b200/) never does.
"""

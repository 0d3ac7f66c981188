"""CPU checkers for the fuzzy-match hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. Nothing under fuzzy_match_b200/ imports it.
"""

"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference NeRF-RPN hot path, used solely as the checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
nerf_rpn_b200/ imports this package.
"""

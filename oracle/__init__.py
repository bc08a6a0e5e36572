"""oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) and from
nowhere else: the product package pytorch_hmm_b200 never imports it.
"""

"""TEST INFRASTRUCTURE ONLY: CPU oracle for the C-matrix hot path.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never from cosmopp_b200/ (the product has no CPU fallback).
"""

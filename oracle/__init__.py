"""CPU oracle for the KLT hot path -- TEST INFRASTRUCTURE, never imported by the product.

Only tests/, __graft_entry__.smoke()/build() and bench.py's CPU-baseline legs may import this package.
"""

"""B200-native batched box-QP engine (host-side mirror of the C-ABI).

Filled in below; see api.py.
"""

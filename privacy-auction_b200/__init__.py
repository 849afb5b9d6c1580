"""privacy-auction_b200 — B200-native engine for the AV-net / NIZK hot path of
Privacy-Auction (SEAL and CCS22 on secp256k1).

The product is `libpa_engine.so` (hand-written sm_100a CUDA behind the C ABI of
include/pa_engine.h) plus the C++ host classes under host/.  This Python package
is only the ctypes plumbing the tests and bench.py use to reach that ABI.
The directory name contains a hyphen, so import it with
`importlib.import_module("privacy-auction_b200")`.
"""
from .engine import Engine, EngineError, load_library, LIB_PATH, SIGNATURES  # noqa: F401

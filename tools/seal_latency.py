"""Wall time of one SEAL auction through pa_seal_run with and without verification, for a few sizes
(development aid; run on a GPU box: python tools/seal_latency.py)."""
import importlib
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pa = importlib.import_module("privacy-auction_b200")
eng = pa.Engine(0)
rnd = random.Random(2024)
for n, c in ((1000, 32), (250, 32), (100, 32), (10, 20)):
    bids = [rnd.randrange(1 << (c - 1)) for _ in range(n)]
    for verify in (True, False):
        best = 1e9
        for rep in range(4):
            t0 = time.perf_counter()
            r = eng.seal_run(11, [n], [c], bids, verify=verify)
            eng.sync()
            best = min(best, time.perf_counter() - t0)
        print(f"n={n:5d} c={c} verify={verify!s:5s}: {best*1e3:8.2f} ms  ({best*1e3/c:.2f} ms per step) ok={all(r['ok'])}")

"""Wall time of one CCS22 auction through pa_ccs22_run (development aid, GPU box)."""
import importlib
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pa = importlib.import_module("privacy-auction_b200")
eng = pa.Engine(0)
rnd = random.Random(7)
for n, c in ((20, 32), (100, 32), (5, 8)):
    bids = [rnd.randrange(1 << (c - 1)) for _ in range(n)]
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        r = eng.ccs22_run(13, [n], [c], [3 % n], bids)
        eng.sync()
        best = min(best, time.perf_counter() - t0)
    print(f"n={n:4d} c={c}: {best*1e3:8.2f} ms ({best*1e3/c:.2f} ms per step) ok={all(v == max(bids) for v in r['max_bid'])}")

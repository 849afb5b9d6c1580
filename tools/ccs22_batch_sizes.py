"""Throughput of pa_ccs22_run against the number of auctions in the lock-step batch (development aid, GPU box)."""
import importlib
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pa = importlib.import_module("privacy-auction_b200")
eng = pa.Engine(0)
for A in (512, 2048, 8192):
    r2 = random.Random(2200)
    bids = [r2.randrange(1 << 31) for _ in range(20 * A)]
    ev = [r2.randrange(20) for _ in range(A)]
    ids = list(range(A))
    eng.ccs22_run(13, [20] * A, [32] * A, ev, bids, auction_ids=ids)
    eng.sync()
    t0 = time.perf_counter()
    rr = eng.ccs22_run(13, [20] * A, [32] * A, ev, bids, auction_ids=ids)
    eng.sync()
    dt = time.perf_counter() - t0
    ok = all(rr["max_bid"][20 * a + i] == max(bids[20 * a:20 * a + 20]) for a in range(A) for i in range(20))
    print(A, f"{dt:.3f} s", f"{A/dt:.0f} auctions/s", ok)

"""Per-kernel CUDA-event breakdown of pa_seal_run for the two SEAL configurations of bench.py
(development aid; run on a GPU box: python tools/seal_profile.py [config4|config5] [repeats])."""
import importlib
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pa = importlib.import_module("privacy-auction_b200")
which = sys.argv[1] if len(sys.argv) > 1 else "config5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = pa.Engine(0)
if which == "config4":
    rnd = random.Random(2024)
    n, c = [1000], [32]
    bids = [rnd.randrange(1 << 31) for _ in range(1000)]
    ids = None
else:
    A = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    r5 = random.Random(5000)
    n = [r5.randint(1, 20) for _ in range(A)]
    c = [r5.randint(1, 32) for _ in range(A)]
    bids = [r5.randrange(1 << min(c[a], 31)) for a in range(A) for _ in range(n[a])]
    ids = list(range(A))
for rep in range(reps):
    eng.profile_begin()
    t0 = time.perf_counter()
    r = eng.seal_run(11, n, c, bids, verify=True, auction_ids=ids)
    eng.sync()
    dt = time.perf_counter() - t0
    ks = eng.profile_end()
    tot = sum(v["total_ms"] for v in ks.values())
    print(f"--- {which} rep {rep}: wall {dt*1e3:.1f} ms, kernels {tot:.1f} ms, launches {sum(v['launches'] for v in ks.values())}, ok {all(r['ok'])}")
    for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["total_ms"])[:14]:
        print(f"   {k:28s} {v['launches']:5d} launches {v['total_ms']:9.2f} ms  avg {v['total_ms']/v['launches']:.3f}")

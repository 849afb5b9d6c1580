"""Timeline of one small auction (development aid): python tools/seal_timeline_small.py n c  with PA_TIMELINE / PA_ENGINE_LIB set"""
import importlib, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pa = importlib.import_module("privacy-auction_b200")
n, c = int(sys.argv[1]), int(sys.argv[2])
eng = pa.Engine(0)
rnd = random.Random(3)
bids = [rnd.randrange(1 << min(c, 31)) for _ in range(n)]
ts = []
for rep in range(5):
    if rep == 4:
        eng.profile_begin()
    t0 = time.perf_counter()
    r = eng.seal_run(7, [n], [c], bids, verify=True)
    eng.sync()
    ts.append(round((time.perf_counter() - t0) * 1e3, 2))
eng.profile_end()
print(n, c, ts, r["ok"])
eng.close()

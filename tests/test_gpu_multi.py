"""Multi-GPU: one SEAL auction sharded by bidder slice over 2 ranks with an NCCL all-gather
of each round's points (BASELINE config 4 in small).  Needs >= 2 GPUs; skipped otherwise."""
import os
import pickle
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, pickle, sys
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
import seal_flow
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
pa = importlib.import_module("privacy-auction_b200")
from importlib import import_module
D = import_module("privacy-auction_b200.distributed")
eng = pa.Engine(int(os.environ["LOCAL_RANK"]))
out = {}
for path in sys.argv[3:]:
    t = seal_flow.parse_transcript(open(path, "rb").read())
    res = D.seal_run_sharded(eng, t["seed"], t["n"], t["c"], t["bids"], verify=True, sections=True)
    out[path] = res
pickle.dump(out, open(os.path.join(sys.argv[2], f"r{rank}.pkl"), "wb"))
dist.destroy_process_group()
'''


def test_sharded_auction_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import seal_flow
    import glob
    # every golden auction: n >= 2 runs the phase-major sharded schedule, n = 1 (rank 1 owns nobody) the step-major one
    golds = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "seal_n*.bin")))
    assert len(golds) >= 7
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29655", str(w), ROOT, str(tmp_path)] + golds, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    parts = [pickle.load(open(tmp_path / f"r{k}.pkl", "rb")) for k in range(2)]
    for path in golds:
        gold = open(path, "rb").read()
        t = seal_flow.parse_transcript(gold)
        n, c = t["n"], t["c"]
        secs = []
        for res in (parts[0][path], parts[1][path]):
            lo, hi = res["slice"]
            m = hi - lo
            assert res["ok_all"] and res["max_bid_all"] == max(t["bids"])
            if m == 0:
                continue                      # more ranks than bidders: this rank only took part in the exchanges
            sec = {"commit": {}, "commit_ok": {}, "r1": [], "r1_ok": [], "r2": [], "r2_ok": [], "r3": []}
            for q in range(m):
                sec["commit"][lo + q] = res["commit"][736 * c * q:736 * c * (q + 1)]
                sec["commit_ok"][lo + q] = all(res["commit_ok"][c * q:c * (q + 1)])
            for step in range(c):
                r1, r1ok, r2, r2ok = {}, {}, {}, {}
                for q in range(m):
                    o = step * m + q
                    r1[lo + q] = res["r1"][320 * o:320 * (o + 1)]
                    r1ok[lo + q] = bool(res["r1_ok"][o])
                    tag = res["r2_tag"][o]
                    r2[lo + q] = int(tag).to_bytes(4, "little") + res["r2_b"][64 * o:64 * (o + 1)] + \
                        res["r2_proof"][1344 * o:1344 * o + (672 if tag == 1 else 1344)]
                    r2ok[lo + q] = bool(res["r2_ok"][o])
                sec["r1"].append(r1); sec["r1_ok"].append(r1ok); sec["r2"].append(r2); sec["r2_ok"].append(r2ok)
                sec["r3"].append(res["r3"][step])
            secs.append(sec)
        assert seal_flow.assemble_transcript(n, c, t["seed"], t["bids"], secs) == gold

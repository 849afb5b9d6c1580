"""Multi-GPU: one SEAL auction sharded by bidder slice over the ranks (BASELINE config 4), both transports:
  xchg  the ranks' kernels exchange their per-step sums through peer windows in each other's HBM
        (pa_xchg_*, CUDA IPC over NVLink) - no host round trip per step, Y reconstruction sharded too;
  nccl  an NCCL all-gather called back from the engine at every exchange.
Every golden auction of the unmodified reference, the full-size config-4 digest of the oracle, and an auction
with forced draw rejections (all ranks must fall back to step-major together) must come out byte-identical.
Needs >= 2 GPUs; skipped otherwise (bench.py repeats the config-4 digest check at every world size)."""
import glob
import hashlib
import json
import os
import pickle
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, json, os, pickle, sys
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
import seal_flow
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
pa = importlib.import_module("privacy-auction_b200")
D = importlib.import_module("privacy-auction_b200.distributed")
eng = pa.Engine(int(os.environ["LOCAL_RANK"]))
D.connect_peer_windows(eng)
out = {}
jobs = json.load(open(sys.argv[3]))
for job in jobs:
    eng.debug_set(1, job.get("reject_bits", 0))
    before = eng.reruns
    res = D.seal_run_sharded(eng, job["seed"], job["n"], job["c"], job["bids"], verify=True, sections=True,
                             transport=job["transport"], schedule=job.get("schedule", 0))
    res["reruns"] = eng.reruns - before
    out[job["key"]] = res
eng.debug_set(1, 0)
pickle.dump(out, open(os.path.join(sys.argv[2], f"r{rank}.pkl"), "wb"))
dist.destroy_process_group()
'''


def _stitch(seal_flow, parts, key, n, c, seed, bids):
    secs = []
    for part in parts:
        res = part[key]
        lo, hi = res["slice"]
        m = hi - lo
        if m == 0:
            continue                      # more ranks than bidders: this rank stayed out
        assert res["ok_all"] and res["max_bid_all"] == max(bids), key
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
    return seal_flow.assemble_transcript(n, c, seed, bids, secs)


def test_sharded_auction_two_gpus(tmp_path, oracle):
    """world = 2 by default; PA_TEST_WORLD=4 or 8 runs the same jobs on more ranks (small auctions then have ranks that
    own nobody and stay out of the exchanges: pa_xchg_skip)."""
    import torch
    world = int(os.environ.get("PA_TEST_WORLD", "2"))
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import seal_flow
    import secp256k1_py as E
    golds = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "seal_n*.bin")))
    assert len(golds) >= 7
    big = json.load(open(os.path.join(ROOT, "tests", "golden", "large_config_digests.json")))["config4_uniform"]
    jobs, want = [], {}
    for transport in ("xchg", "nccl"):
        # every golden auction: n >= 2 runs the phase-major sharded schedule, n = 1 (rank 1 owns nobody) the step-major one
        for path in golds:
            t = seal_flow.parse_transcript(open(path, "rb").read())
            key = f"{transport}:{os.path.basename(path)}"
            jobs.append(dict(key=key, transport=transport, seed=t["seed"], n=t["n"], c=t["c"], bids=t["bids"]))
            want[key] = open(path, "rb").read()
        # the step-major schedule over the same transport
        t = seal_flow.parse_transcript(open(golds[3], "rb").read())
        key = f"{transport}:step-major"
        jobs.append(dict(key=key, transport=transport, seed=t["seed"], n=t["n"], c=t["c"], bids=t["bids"], schedule=1))
        want[key] = open(golds[3], "rb").read()
        # forced draw rejections: every rank must fall back to step-major, together
        key = f"{transport}:reject"
        jobs.append(dict(key=key, transport=transport, seed=31, n=6, c=7, bids=[5, 100, 77, 0, 100, 64], reject_bits=4))
        E.REJECT_BITS = 4
        try:
            want[key] = seal_flow.SealFlow(oracle, 6, 7, 31, [5, 100, 77, 0, 100, 64]).run()
        finally:
            E.REJECT_BITS = 0
        # BASELINE config 4 at full size against the oracle's digest
        key = f"{transport}:config4"
        jobs.append(dict(key=key, transport=transport, seed=big["seed"], n=big["n"], c=big["c"], bids=big["bids"]))
    (tmp_path / "jobs.json").write_text(json.dumps(jobs))
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29655", str(w), ROOT, str(tmp_path), str(tmp_path / "jobs.json")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    parts = [pickle.load(open(tmp_path / f"r{k}.pkl", "rb")) for k in range(world)]
    for job in jobs:
        key = job["key"]
        got = _stitch(seal_flow, parts, key, job["n"], job["c"], job["seed"], job["bids"])
        if key.endswith(":config4"):
            assert hashlib.sha256(got).hexdigest() == big["sha256"], key
        else:
            assert got == want[key], key
        if key.endswith(":reject"):
            # the phase-major schedule notices the rejected draws and runs the auction again; with call-backs and a rank that
            # owns nobody the schedule is step-major from the start, which carries the counters sequentially (no rerun)
            slice_ = (job["n"] + world - 1) // world
            expect = 1 if job["transport"] == "xchg" or (world - 1) * slice_ < job["n"] else 0
            assert all(p[key]["reruns"] == expect for p in parts if p[key]["slice"][1] > p[key]["slice"][0]), key   # ranks that own bidders
        assert parts[0][key]["transport"] == job["transport"]

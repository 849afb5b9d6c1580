#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generates tests/golden/large_config_digests.json (build container only).

BASELINE.json configs 4 and 5 at the sizes the metric is quoted on are beyond the unmodified
reference on CPU (about 10 core-days), so they are pinned with the Tier-B
restatement (oracle/pa_oracle.c on libcrypto + tests/seal_flow.py), which itself reproduces every
Tier-A golden transcript byte for byte (tests/test_oracle_golden.py):

  config 4   ONE auction, n = 1000 bidders x c = 32 bits, every proof verified once, played by
             W processes over gloo (each a contiguous bidder slice, all-gather of X and b only,
             exactly tests/test_sharded_gloo.py), three bid vectors:
               uniform  random.Random(2024) bids in [0, 2^31), seed 4 (bench.py's auction; junction at step 1)
               late     bids in [0, 2^20): the first deciding step is >= 11 (long speculative stage-1 window)
               zero     all bids 0: no deciding step at all, 32 stage-1 rounds
  config 5   64 genTests-shaped auctions, n ~ U{1..20}, c ~ U{1..32} (reference SEAL/tests/genTests.py:15-16),
             PA stream (seed, auction << 32 | bidder) as a lock-step batch of pa_seal_run draws them.

Only SHA-256 digests are committed (a transcript of config 4 is 79 MB).  Follows
/root/reference/SEAL/bidder.cpp:1271-1421 and SEAL/main.cpp:65-120 through seal_flow.SealFlow.

usage:  python tests/golden/make_large_digests.py [--world 8] [--only uniform,late,zero,config5]
"""
import argparse
import hashlib
import json
import os
import pickle
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "large_config_digests.json")
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def config4_vectors():
    n, c = 1000, 32
    r = random.Random(2024)
    uniform = [r.randrange(1 << 31) for _ in range(n)]
    r = random.Random(2025)
    late = [r.randrange(1 << 20) for _ in range(n)]
    return {
        "uniform": {"n": n, "c": c, "seed": 4, "bids": uniform, "bids_from": "random.Random(2024).randrange(1 << 31) x 1000"},
        "late": {"n": n, "c": c, "seed": 5, "bids": late, "bids_from": "random.Random(2025).randrange(1 << 20) x 1000"},
        "zero": {"n": n, "c": c, "seed": 6, "bids": [0] * n, "bids_from": "all zero"},
    }


def config5_sample(seed=11, count=64, rseed=5):
    """n ~ U{1..20}, c ~ U{1..32} like genTests.py:15-16; bids uniform below 2^min(c, 31)."""
    r = random.Random(rseed)
    n = [r.randint(1, 20) for _ in range(count)]
    c = [r.randint(1, 32) for _ in range(count)]
    # make sure the corners the reference's generator can emit are in the sample
    n[0], c[0] = 20, 32
    n[1], c[1] = 1, 32
    n[2], c[2] = 20, 1
    n[3], c[3] = 1, 1
    bids = [[r.randrange(1 << min(c[a], 31)) for _ in range(n[a])] for a in range(count)]
    bids[4] = [0] * n[4]
    return {"seed": seed, "n": n, "c": c, "bids": bids}


def _worker4(rank, world, port, spec, outdir):
    import torch
    import torch.distributed as dist
    import oracle_lib
    import seal_flow
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, c = spec["n"], spec["c"]
    slice_ = (n + world - 1) // world
    lo, hi = min(n, rank * slice_), min(n, (rank + 1) * slice_)

    def exchange(kind, local):
        send = torch.zeros(slice_ * 64, dtype=torch.uint8)
        send[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8)
        recv = torch.empty(world * slice_ * 64, dtype=torch.uint8)
        dist.all_gather_into_tensor(recv, send)
        return bytes(recv.numpy().tobytes()[:n * 64])

    fl = seal_flow.SealFlow(oracle_lib.Oracle(), n, c, spec["seed"], spec["bids"], mine=range(lo, hi), exchange=exchange)
    sec = fl.run_sections()
    pickle.dump((sec, fl.max_bid), open(os.path.join(outdir, f"r{rank}.pkl"), "wb"))
    dist.destroy_process_group()


def run_config4(name, spec, world, port):
    import torch.multiprocessing as mp
    import seal_flow
    t0 = time.time()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker4, args=(world, port, spec, td), nprocs=world, join=True)
        parts = [pickle.load(open(os.path.join(td, f"r{r}.pkl"), "rb")) for r in range(world)]
    n, c = spec["n"], spec["c"]
    out = seal_flow.assemble_transcript(n, c, spec["seed"], spec["bids"], [p[0] for p in parts])
    assert all(p[1] == max(spec["bids"]) for p in parts), "Tier-B max bid wrong"
    assert seal_flow.transcript_ok(out), "Tier-B: a proof failed verification"
    r3 = parts[0][0]["r3"]
    d = dict(spec)
    d.update({"protocol": "seal", "bytes": len(out), "sha256": hashlib.sha256(out).hexdigest(), "r3": r3,
              "first_deciding_step": next((s for s in range(c) if r3[s]), None),
              "oracle": f"Tier B (oracle/pa_oracle.c + tests/seal_flow.py), {world} gloo processes, every proof verified once",
              "oracle_seconds": round(time.time() - t0, 1)})
    # section digests, to localise a mismatch: the bytes of each step as they appear in the transcript
    t = seal_flow.parse_transcript(out)
    d["commit_sha256"] = hashlib.sha256(b"".join(b"".join(row) for row in t["commit"])).hexdigest()
    # per bidder: everything that bidder published, in transcript order - lets the ranks of a sharded run check
    # their own slices (bench.py) without moving 78 MB of records to one place
    d["bidder_sha256"] = [bidder_digest(t, j) for j in range(n)]
    d["step_sha256"] = [hashlib.sha256(b"".join(s["r1"]) + b"".join(int(tag).to_bytes(4, "little") + b + pr for tag, b, pr in s["r2"])).hexdigest()
                        for s in t["steps"]]
    return d


def bidder_digest(t, j):
    """SHA-256 of what bidder j published: its c commitment records, then per step its round-one record,
    LE32 stage tag, cryptogram and proof (t = seal_flow.parse_transcript(...))."""
    h = hashlib.sha256(b"".join(t["commit"][j]))
    for s in t["steps"]:
        tag, b, pr = s["r2"][j]
        h.update(s["r1"][j] + int(tag).to_bytes(4, "little") + b + pr)
    return h.hexdigest()


def _one5(args):
    import oracle_lib
    import seal_flow
    seed, a, n, c, bids = args
    fl = seal_flow.SealFlow(oracle_lib.Oracle(), n, c, seed, bids, auction=a)
    out = fl.run()
    assert fl.ok
    return hashlib.sha256(out).hexdigest(), len(out)


def run_config5(world):
    import multiprocessing as mp
    s = config5_sample()
    t0 = time.time()
    with mp.get_context("spawn").Pool(world) as pool:
        res = pool.map(_one5, [(s["seed"], a, s["n"][a], s["c"][a], s["bids"][a]) for a in range(len(s["n"]))], chunksize=1)
    s.update({"protocol": "seal", "sha256": [r[0] for r in res], "bytes": [r[1] for r in res],
              "oracle": "Tier B, one process per auction, PA stream (seed, auction << 32 | bidder)",
              "oracle_seconds": round(time.time() - t0, 1)})
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--only", default="uniform,late,zero,config5")
    args = ap.parse_args()
    D = json.load(open(OUT)) if os.path.exists(OUT) else {}
    port = 29700
    for name, spec in config4_vectors().items():
        if name not in args.only.split(","):
            continue
        port += 1
        D[f"config4_{name}"] = run_config4(name, spec, args.world, port)
        json.dump(D, open(OUT, "w"), indent=1)
        print(name, D[f"config4_{name}"]["sha256"], D[f"config4_{name}"]["oracle_seconds"], "s", flush=True)
    if "config5" in args.only.split(","):
        D["config5_sample"] = run_config5(args.world)
        json.dump(D, open(OUT, "w"), indent=1)
        print("config5", D["config5_sample"]["oracle_seconds"], "s", flush=True)


if __name__ == "__main__":
    main()

"""The bidder-slice partitioning of ONE auction over several processes (SURVEY.md
section 8e, BASELINE config 4) on CPU: world_size 2 and 3 over gloo, each process playing a
contiguous id range with the libcrypto oracle as backend, exchanging only the X of round one
and the b of round two by all-gather (padded to equal slices, as pa_seal_run does).  The
stitched transcript must equal the single-process one and the reference's golden."""
import os
import pickle
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "seal_n5_c5_s11.bin")
GOLD2 = os.path.join(ROOT, "tests", "golden", "seal_n4_c6_s7.bin")


def _worker(rank, world, port, path, outdir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle_lib
    import seal_flow
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = seal_flow.parse_transcript(open(path, "rb").read())
    n = t["n"]
    slice_ = (n + world - 1) // world
    lo, hi = min(n, rank * slice_), min(n, (rank + 1) * slice_)

    def exchange(kind, local):
        send = torch.zeros(slice_ * 64, dtype=torch.uint8)
        send[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8)
        recv = torch.empty(world * slice_ * 64, dtype=torch.uint8)
        dist.all_gather_into_tensor(recv, send)
        return bytes(recv.numpy().tobytes()[:n * 64])

    fl = seal_flow.SealFlow(oracle_lib.Oracle(), n, t["c"], t["seed"], t["bids"], mine=range(lo, hi), exchange=exchange)
    sec = fl.run_sections()
    # verdict reduction: one MIN all-reduce of a single word
    ok = torch.tensor([int(all(sec["commit_ok"].values()) and all(all(d.values()) for d in sec["r1_ok"] + sec["r2_ok"]))])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    pickle.dump((sec, fl.max_bid, int(ok)), open(os.path.join(outdir, f"r{rank}.pkl"), "wb"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,path,port", [(2, GOLD, 29611), (3, GOLD, 29612), (2, GOLD2, 29613)])
def test_sharded_auction_equals_reference(tmp_path, world, path, port):
    import seal_flow
    mp.spawn(_worker, args=(world, port, path, str(tmp_path)), nprocs=world, join=True)
    gold = open(path, "rb").read()
    t = seal_flow.parse_transcript(gold)
    parts = [pickle.load(open(tmp_path / f"r{r}.pkl", "rb")) for r in range(world)]
    assert all(p[1] == max(t["bids"]) and p[2] == 1 for p in parts)
    assert seal_flow.assemble_transcript(t["n"], t["c"], t["seed"], t["bids"], [p[0] for p in parts]) == gold

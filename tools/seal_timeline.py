"""Device-side timeline of ONE n = 1000 x 32-bit SEAL auction (BASELINE config 4), sharded over the ranks it is launched
with (development aid; run on a GPU box):
    PA_TIMELINE=gpurun_out/tl python tools/seal_timeline.py                      # one GPU
    PA_TIMELINE=gpurun_out/tl python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/seal_timeline.py
Every rank appends "start_ms dur_ms lane kernel" per launch of the last (profiled) run to $PA_TIMELINE.<device>; rank 0
prints the wall times.  Gaps between kernels on the main lane are host work, synchronisation and exchange latency."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

pa = importlib.import_module("privacy-auction_b200")
D = importlib.import_module("privacy-auction_b200.distributed")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = pa.Engine(local)
if world > 1:
    D.connect_peer_windows(eng)
G = json.load(open(os.path.join(ROOT, "tests", "golden", "large_config_digests.json")))[sys.argv[1] if len(sys.argv) > 1 else "config4_uniform"]
n, c, seed, bids = G["n"], G["c"], G["seed"], G["bids"]


def run():
    if world == 1:
        return eng.seal_run(seed, [n], [c], bids, verify=True)
    return D.seal_run_sharded(eng, seed, n, c, bids, verify=True)


def sync():
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    eng.sync()


times = []
for rep in range(4):
    sync()
    if rep == 3:
        eng.profile_begin()
    t0 = time.perf_counter()
    r = run()
    eng.sync()
    times.append((time.perf_counter() - t0) * 1e3)
ks = eng.profile_end()
sync()
if rank == 0:
    print(f"world {world}: wall ms per run {[round(t, 2) for t in times]}; kernel ms of the last run {round(sum(v['total_ms'] for v in ks.values()), 2)}")
eng.close()
if dist:
    dist.destroy_process_group()

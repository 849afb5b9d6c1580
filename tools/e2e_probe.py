"""What limits the end-to-end figure (host buffers in, host buffers out) when 8 ranks share one host?  (development aid)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_probe.py
Every rank runs the bench's e2e step (pa_fixed_base_mul + pa_var_base_mul on pinned host buffers of 2^20 items) in three
arrangements: all ranks in lock step (what bench.py times); odd ranks with the two calls swapped; every rank delayed by
rank/world of a step.  Plus the copy bandwidth each rank gets when all ranks copy at once (H2D and D2H together)."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

pa = importlib.import_module("privacy-auction_b200")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = pa.Engine(local)
n = 1 << 20
rng = np.random.default_rng(1234 + rank)
h_kf = torch.from_numpy(np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).copy()).pin_memory()
h_kv = torch.from_numpy(np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).copy()).pin_memory()
h_out_f = torch.empty(64 * n, dtype=torch.uint8).pin_memory()
h_out_v = torch.empty(64 * n, dtype=torch.uint8).pin_memory()
eng._check(eng.lib.pa_fixed_base_mul(eng.ctx, h_kf.data_ptr(), h_out_f.data_ptr(), n))
h_bases = h_out_f.clone().pin_memory()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def fixed():
    eng._check(eng.lib.pa_fixed_base_mul(eng.ctx, h_kf.data_ptr(), h_out_f.data_ptr(), n))


def var():
    eng._check(eng.lib.pa_var_base_mul(eng.ctx, h_bases.data_ptr(), h_kv.data_ptr(), h_out_v.data_ptr(), n))


def run(name, order, delay_s=0.0, steps=10):
    for _ in range(2):
        order[0](); order[1]()
    barrier()
    t0 = time.perf_counter()
    if delay_s:
        time.sleep(delay_s)
    for _ in range(steps):
        order[0](); order[1]()
    mine = time.perf_counter() - t0 - delay_s
    barrier()
    t = torch.tensor([mine], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:34s} slowest rank {float(t.item()) / steps * 1e3:7.2f} ms per step -> {2.0 * n * world * steps / float(t.item()) / 1e6:8.1f} M/s aggregate", flush=True)


run("lock step (bench.py)", (fixed, var))
run("odd ranks swapped", (var, fixed) if rank & 1 else (fixed, var))
run("rank r delayed by r/world steps", (fixed, var), delay_s=0.021 * rank / world)
run("fixed-base only x2", (fixed, fixed))
run("variable-base only x2", (var, var))

# copy bandwidth with every rank copying at once
d_in = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
h_in = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for which in ("h2d", "d2h", "both"):
    barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        if which in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if which in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = (8 * 128 * (2 if which == "both" else 1)) / 1024 / dt
    t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"copy {which:5s}: slowest rank {float(t.item()):6.1f} GiB/s with {world} ranks copying at once", flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()

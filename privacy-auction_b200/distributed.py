"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL).

The hot path partitions two ways:
  * independent auctions / independent batches of scalar mults: every rank works on
    its own share, there is NO data-path collective;
  * ONE auction sharded by bidder slice: the X_i of round one (64 B per bidder and step) are
    all-gathered once for all steps, per step every rank contributes the sum of its
    cryptograms (128 B), and one word of verdict is MIN-reduced at the end.  (The step-major
    schedule, used when some rank would own no bidder, gathers X_i and b_i once per step.)
    The engine (pa_seal_run) calls back into `allgather` below when the send buffer is ready.
"""
import torch
import torch.distributed as dist


def connect_peer_windows(engine):
    """Create this rank's peer exchange window and map everybody else's (CUDA IPC over NVLink): after this the
    kernels of a sharded auction exchange their per-step sums by writing into each other's HBM.  The 64-byte
    handles travel through torch.distributed once; nothing else does afterwards.
    Returns True when EVERY rank has mapped every window; if any rank could not (no peer access between two GPUs,
    IPC refused in a container), all ranks close their windows and return False together, and seal_run_sharded's
    transport "auto" then uses the NCCL call-backs."""
    world = dist.get_world_size()
    if engine.xchg_world == world:
        return True
    rank = dist.get_rank()
    ok = 1
    try:
        mine = engine.xchg_create()
    except Exception:
        mine, ok = b"", 0
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    if ok and all(len(h) == 64 for h in handles):
        try:
            engine.xchg_connect(handles, rank)
        except Exception:
            ok = 0
    else:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        try:
            engine.xchg_close()
        except Exception:
            pass
        engine.xchg_world = 0
        return False
    return True


def seal_run_sharded(engine, seed, n, c, bids_all, verify=True, sections=False, transport="auto", schedule=0):
    """Run ONE SEAL auction of n bidders sharded over the ranks of the default process group.
    Every rank passes the same (seed, n, c, bids_all).  Returns the engine's result dict for
    the local slice plus 'ok_all' (AND over ranks), 'max_bid_all' and 'slice' = (lo, hi).
    transport: "xchg" = kernels exchange through the peer windows (connect_peer_windows), no host round
    trip per step; "nccl" = an NCCL all-gather called back from the engine per exchange; "auto" = xchg when
    the windows are connected."""
    rank, world = dist.get_rank(), dist.get_world_size()
    slice_ = (n + world - 1) // world
    lo, hi = min(n, rank * slice_), min(n, (rank + 1) * slice_)
    dev = torch.device("cuda", torch.cuda.current_device())
    if transport == "auto":
        transport = "xchg" if engine.xchg_world == world else "nccl"
    if transport == "xchg":
        assert engine.xchg_world == world, "connect_peer_windows(engine) first"
        if hi > lo:
            res = engine.seal_run(seed, [n], [c], list(bids_all[lo:hi]), verify=verify, sections=sections, schedule=schedule,
                                  shard=dict(lo=lo, hi=hi, slice=slice_, use_xchg=True))
            res["max_bid_all"] = res["max_bid"][0]   # every rank derives it from the same all-rank sums
        else:   # more ranks than bidders: this rank owns nobody and stays out of the exchanges
            engine.xchg_skip()
            res = {"max_bid": [0], "ok": [True], "ok_all": None, "max_bid_all": None}
        res["slice"] = (lo, hi)
        res["transport"] = "xchg"
        return res
    cap = max(c * slice_ * 64, 128)
    send = torch.zeros(cap, dtype=torch.uint8, device=dev)
    recv = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    # the phase-major schedule needs every rank to own bidders (all ranks take the same decisions)
    phase_major = (world - 1) * slice_ < n and schedule != 1
    sizes = {0: slice_ * 64, 1: slice_ * 64, 2: c * slice_ * 64, 3: 128}

    def allgather(which):
        k = sizes[which]
        dist.all_gather_into_tensor(recv[:world * k], send[:k])   # NCCL over NVLink
        # wait for the collective only: a device-wide synchronize would also wait for the engine's side
        # lanes (proofs / verification of earlier steps) and serialise them with the exchange
        torch.cuda.current_stream().synchronize()
        return 0

    if hi > lo:
        res = engine.seal_run(seed, [n], [c], list(bids_all[lo:hi]), verify=verify, sections=sections, schedule=1 if not phase_major else schedule,
                              shard=dict(lo=lo, hi=hi, slice=slice_, d_send=send.data_ptr(), d_recv=recv.data_ptr(), allgather=allgather,
                                         xchg_bytes=cap if phase_major else 0))
    else:  # more ranks than bidders: still take part in the exchanges
        for _ in range(2 * c):
            allgather(0)
        res = {"max_bid": [0], "ok": [True]}
    ok = torch.tensor([1 if all(res["ok"]) else 0], dtype=torch.int32, device=dev)
    mb = torch.tensor([res["max_bid"][0]], dtype=torch.int64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.all_reduce(mb, op=dist.ReduceOp.MAX)
    res["ok_all"] = bool(ok.item())
    res["max_bid_all"] = int(mb.item())
    res["slice"] = (lo, hi)
    res["transport"] = "nccl"
    return res

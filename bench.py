#!/usr/bin/env python
"""bench.py — headline benchmark of the engine (driver contract in the task spec).

Workload (BASELINE.json configs[2], the 1-GPU configuration the metric
"EC scalar-mults/s" is quoted on): per GPU and per step, 2^20 fixed-base
scalar multiplications g^k and 2^20 variable-base scalar multiplications P^k on
secp256k1, seeded synthetic scalars, variable bases P = g^k' produced by the
fixed-base kernel.  One step = one pass over that batch.

  value      : scalar mults / s, inputs resident in HBM, CUDA-event timed on the
               engine's stream, max over ranks, aggregate over all GPUs (weak scaling)
  e2e        : the same batch through the host-buffer C-ABI calls
               (pa_fixed_base_mul / pa_var_base_mul) from pinned host memory,
               H2D and D2H copies inside the timed region
  roofline   : integer-pipe (IMAD) roofline of the dominant kernel k_var_base,
               algorithmic work from SURVEY.md §8(d) (2,900 field mults x 272
               IMAD units per variable-base mult), duration from CUDA events
               bracketing the kernel inside the timed region
  cpu_baseline / --impl reference : OpenSSL libcrypto EC_POINT_mul in the
               reference's call shapes (oracle/_ref/ecmul_ref) on the host cores
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_KIND = 1 << 20
METRIC = "ec_scalar_mults_per_s"
UNIT = "scalar-mults/s"
WORKLOAD = "configs[2]: batched EC microbench, 2^20 fixed-base + 2^20 variable-base scalar mults on secp256k1 per GPU per step"
# SURVEY.md §8(d) algorithmic work figures (nominal double-and-add / comb on 32-bit IMAD)
FM_VAR, FM_FIXED, IMAD_PER_FM = 2900, 712, 272
# Executed 32x32->64 multiply-adds (SASS IMAD.WIDE) per scalar multiplication, counted by ncu on the
# shipped kernels (profiles/r01f_opcode_mix.txt, both kernels from the same capture)
WIDE_VAR, WIDE_FIXED = 100233, 11299
ECMUL_REF = os.path.join(ROOT, "oracle", "_ref", "ecmul_ref")


def pin_to_gpu_numa_node(gpu_index):
    """Several ranks on one host: keep this process (and the pinned host buffers it is about to allocate, first
    touch) on the CPU cores NVML reports as local to its GPU, so that the H2D / D2H copies of the end-to-end
    figure do not cross sockets.  Returns the number of cores it was pinned to (0: left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = gpu_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if gpu_index < len(ids) and ids[gpu_index].isdigit():
                phys = int(ids[gpu_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_ecmul_ref(threads, pairs, seed=1):
    """The reference arm: libcrypto EC_POINT_mul, `threads` independent threads."""
    if os.path.exists(ECMUL_REF):
        out = subprocess.run([ECMUL_REF, str(threads), str(pairs), str(seed)], capture_output=True, text=True, check=True).stdout
        r = json.loads(out.strip().splitlines()[-1])
        return {"kind": "reference", "mults": r["mults"], "seconds": r["seconds"], "threads": threads}
    # prebuilt binary missing: time the oracle port (single thread) instead
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import random
    ora = oracle_lib.Oracle()
    rnd = random.Random(seed)
    ks = b"".join(rnd.getrandbits(256).to_bytes(32, "big") for _ in range(pairs))
    t0 = time.perf_counter()
    pts = ora.fixed_base_mul(ks)
    ora.var_base_mul(pts, ks[::-1])
    return {"kind": "port", "mults": 2 * pairs, "seconds": time.perf_counter() - t0, "threads": 1}


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed
    region (NVML every 20 ms; nvidia-smi every 200 ms if NVML is unavailable)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag = gpu_index, threading.Event()
        self.sm, self.mx, self.power, self.reasons, self.how = [], [], [], set(), "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # torch honours CUDA_VISIBLE_DEVICES; NVML does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv, self.how = None, "nvidia-smi"

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        self.sm.append(float(r[1]))
        self.mx.append(float(r[2]))
        for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
            if val.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nv else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nv else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "power_w_max": max(self.power) if self.power else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "how": self.how}


def point_stream_roofline(kstats, n):
    """HBM roofline of the one point-stream kernel (north_star: HBM GB/s only for table / point-stream
    kernels): k_normalize reads a 96-B Jacobian triple twice (Z, then X Y Z), writes and re-reads a
    32-B prefix product and writes the 64-B affine point = 320 B per point."""
    nz = kstats.get("k_normalize")
    if not nz:
        return None
    ms = nz["total_ms"] / max(nz["launches"], 1)
    peak, src = 6650.0, "fallback of B200_PROFILING.md"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    achieved = 320.0 * n / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "k_normalize", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": src, "avg_launch_ms": ms,
            "note": "not HBM-bound: one 270-multiplication field inversion per 16 points dominates (integer pipe)"}


def proof_throughput(eng, torch, n=1 << 15, seed=77):
    """proof-verifies/s and proofs/s per NIZK kind at a batch large enough to fill the GPU
    (n proofs per kind, device-resident, CUDA events on the engine's stream).  Statements are
    valid SEAL statements built with the engine itself; every verdict must be 1."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sc = lambda k=1: rng.integers(0, 256, size=(n, 32 * k), dtype=np.uint8)
    cat = lambda *a: np.ascontiguousarray(np.concatenate(a, axis=1))
    P = lambda b: np.frombuffer(b, dtype=np.uint8).reshape(n, -1)
    x, r, al, be, xj, rj = sc(), sc(), sc(), sc(), sc(), sc()
    bit = rng.integers(0, 2, size=n, dtype=np.uint8)
    bj = rng.integers(0, 2, size=n, dtype=np.uint8)
    bi = bit & bj
    ids = [int(v) for v in rng.integers(0, 1 << 20, size=n)]
    B = lambda a: np.ascontiguousarray(a).tobytes()   # the ctypes binding takes flat byte strings
    X, R, Y = P(eng.fixed_base_mul(B(x))), P(eng.fixed_base_mul(B(r))), P(eng.fixed_base_mul(B(sc())))
    Xj, Rj, Yj = P(eng.fixed_base_mul(B(xj))), P(eng.fixed_base_mul(B(rj))), P(eng.fixed_base_mul(B(sc())))
    pick = lambda m, a, b: np.where(m[:, None].astype(bool), a, b)
    c1 = P(eng.commit_points(B(al), B(be), B(bit)))
    b1 = P(eng.var_base_mul(B(pick(bit, R, Y)), B(x)))
    c2 = P(eng.commit_points(B(al), B(be), B(bi)))
    Bi = P(eng.var_base_mul(B(pick(bi, R, Y)), B(x)))
    Bj = P(eng.var_base_mul(B(pick(bj, Rj, Yj)), B(xj)))
    cases = {
        "pok": dict(stmt=X, sec=x, args=(), rnd=sc(1), rec=96, mults=2),
        "com": dict(stmt=c1, sec=al, args=(bit,), rnd=sc(3), rec=352, mults=8),
        "s1": dict(stmt=cat(b1, X, Y, R, c1), sec=cat(x, al), args=(bit,), rnd=sc(5), rec=672, mults=16),
        "s2": dict(stmt=cat(Bi, X, R, Bj, Xj, Rj, c2, Y, Yj), sec=cat(x, xj, al), args=(bi, bj), rnd=sc(11), rec=1344, mults=32),
    }
    names = {"pok": "pokdlog", "com": "powfcom", "s1": "stage1", "s2": "stage2"}
    stream = torch.cuda.ExternalStream(eng.stream)
    t_ids = torch.tensor(ids, dtype=torch.int64, device="cuda")
    out = {}
    for kind, cs in cases.items():
        dev = lambda a: torch.from_numpy(np.array(a, copy=True)).cuda()
        d_stmt, d_sec, d_rnd = dev(cs["stmt"]), dev(cs["sec"]), dev(cs["rnd"])
        d_args = [dev(a) for a in cs["args"]]
        d_proofs = torch.empty(n * cs["rec"], dtype=torch.uint8, device="cuda")
        d_verdict = torch.zeros(n, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        prove = getattr(eng.lib, f"pa_{names[kind]}_prove_dev")
        verify = getattr(eng.lib, f"pa_{names[kind]}_verify_dev")
        pargs = [d_stmt.data_ptr(), d_sec.data_ptr()] + [a.data_ptr() for a in d_args] + [t_ids.data_ptr(), d_rnd.data_ptr(), d_proofs.data_ptr(), n]
        vargs = [d_proofs.data_ptr(), d_stmt.data_ptr(), t_ids.data_ptr(), d_verdict.data_ptr(), n]
        times = {}
        for what, fn, a in (("prove", prove, pargs), ("verify", verify, vargs)):
            eng._check(fn(eng.ctx, *a))  # warm-up
            eng.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                eng._check(fn(eng.ctx, *a))
            e1.record(stream)
            eng.sync()
            times[what] = e0.elapsed_time(e1) / 3
        assert bool((d_verdict == 1).all()), f"{kind}: a valid proof was rejected"
        out[kind] = {"batch": n, "verify_ms": times["verify"], "verifies_per_s": n / (times["verify"] * 1e-3),
                     "prove_ms": times["prove"], "proofs_per_s": n / (times["prove"] * 1e-3),
                     "verify_scalar_mults_per_s": cs["mults"] * n / (times["verify"] * 1e-3)}
    return out


def seal_figures(eng, pa, rank, world, dist, torch, config5_auctions=4096):
    """Secondary figures (not the headline `value`): SEAL auctions through pa_seal_run.
      config4: ONE auction, n = 1000 bidders x 32-bit bids, sharded by bidder slice over the ranks
               (one NCCL all-gather of the X of all steps, then 128 B of partial sums per rank and step), every proof verified once;
      config5: a lock-step batch of independent genTests-style auctions (n ~ U{1..20}, c ~ U{1..32},
               reference tests/genTests.py:15-16) per rank, no exchange;
      verifies/s per proof kind from the CUDA-event time of the verify kernels inside the config4 run."""
    import importlib
    import random
    D = importlib.import_module("privacy-auction_b200.distributed")
    out = {}
    rnd = random.Random(2024)
    n4, c4 = 1000, 32
    bids = [rnd.randrange(1 << 31) for _ in range(n4)]
    sync = lambda: (dist.barrier() if dist else None, torch.cuda.synchronize(), eng.sync())

    def run4():
        if world == 1:
            return eng.seal_run(7, [n4], [c4], bids, verify=True)
        r = D.seal_run_sharded(eng, 7, n4, c4, bids, verify=True)
        return {"ok": [r["ok_all"]], "max_bid": [r["max_bid_all"]]}

    run4()  # warm-up (arena growth, module load)
    sync()
    eng.profile_begin()
    t0 = time.perf_counter()
    r = run4()
    sync()
    dt = time.perf_counter() - t0
    ks = eng.profile_end()
    assert r["ok"] == [True] and r["max_bid"] == [max(bids)], "SEAL n=1000 run failed its own checks"
    t_dt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t_dt, op=dist.ReduceOp.MAX)
    dt = float(t_dt.item())
    out["config4_seal_n1000_c32"] = {"auctions_per_s": 1.0 / dt, "seconds": dt, "bidders": n4, "bits": c4,
                                     "partition": "single GPU" if world == 1 else f"bidder slices over {world} GPUs, one NCCL all-gather of the X of all steps, then 128 B of partial sums per rank and step",
                                     "verification": "every proof once (the reference repeats each check n-1 times)"}
    # proofs verified on this rank during the run, per kind
    m = n4 if world == 1 else (min(n4, (rank + 1) * ((n4 + world - 1) // world)) - min(n4, rank * ((n4 + world - 1) // world)))
    decided = sum(1 for s in range(c4) if (max(bids) >> (c4 - 1 - s)) & 1)
    first = next(s for s in range(c4) if (max(bids) >> (c4 - 1 - s)) & 1)
    n_s1, n_s2 = m * (first + 1), m * (c4 - first - 1)
    counts = {"pok": 2 * m * c4 + 2 * m * c4, "com": m * c4, "s1": n_s1, "s2": n_s2}
    ver = {}
    for kind, cnt in counts.items():
        ms = sum(v["total_ms"] for k, v in ks.items() if k in (f"k_verify_derive<{kind}>", f"k_verify_checks<{kind}>"))
        if ms > 0 and cnt > 0:
            ver[kind] = {"proofs": cnt, "kernel_ms": ms, "verifies_per_s_per_gpu": cnt / (ms * 1e-3)}
    out["proof_verifies_inside_n1000_auction"] = ver
    if rank == 0:
        out["proof_throughput_large_batch"] = proof_throughput(eng, torch)
    out["config4_kernels_ms"] = {k: round(v["total_ms"], 3) for k, v in ks.items()}

    # configs 1 and 2 as ONE auction through the runners (latency, rank 0 only)
    if rank == 0:
        r1 = random.Random(1)
        b1 = [r1.randrange(1 << 20) for _ in range(10)]
        b2 = [r1.randrange(1 << 31) for _ in range(20)]
        for _ in range(2):
            eng.sync()
            t0 = time.perf_counter()
            ra = eng.seal_run(1, [10], [20], b1, verify=True)
            eng.sync()
            dt1 = time.perf_counter() - t0
            t0 = time.perf_counter()
            rap = eng.seal_run(1, [10], [20], b1, verify=9)   # every proof verified n - 1 = 9 times: the reference's work
            eng.sync()
            dt1p = time.perf_counter() - t0
            t0 = time.perf_counter()
            rb = eng.ccs22_run(2, [20], [32], [7], b2)
            eng.sync()
            dt2c = time.perf_counter() - t0
        assert ra["ok"] == [True] and ra["max_bid"] == [max(b1)] and all(v == max(b2) for v in rb["max_bid"])
        assert rap["ok"] == [True] and rap["max_bid"] == [max(b1)]
        out["config1_seal_n10_c20_one_auction"] = {"seconds": dt1, "seconds_all_pairs_work": dt1p,
                                                   "path": "pa_seal_run, phase-major schedule, every proof verified once (all_pairs_work: n - 1 = 9 times each, what the reference's 10 bidders do)",
                                                   "reference": "./SEAL 10 20: 51.6-60.6 s on one core (all-pairs verification); per-party CLI on the engine: 2.6 s"}
        out["config2_ccs22_n20_c32_one_auction"] = {"seconds": dt2c, "path": "pa_ccs22_run, phase-major schedule (step-major: 0.207 s)",
                                                    "reference": "./CCS22 20 32: 4.46 s on one core; per-party CLI on the engine: 2.7 s"}

    # config 5 sample: independent auctions, each rank its own batch
    A = config5_auctions
    r5 = random.Random(5000 + rank)
    n5 = [r5.randint(1, 20) for _ in range(A)]
    c5 = [r5.randint(1, 32) for _ in range(A)]
    b5 = [r5.randrange(1 << min(c5[a], 31)) for a in range(A) for _ in range(n5[a])]
    ids5 = [rank * A + a for a in range(A)]
    eng.seal_run(11, n5, c5, b5, verify=True, auction_ids=ids5)   # warm-up (device pools, arenas)
    sync()
    t0 = time.perf_counter()
    r = eng.seal_run(11, n5, c5, b5, verify=True, auction_ids=ids5)
    sync()
    dt5 = time.perf_counter() - t0
    off, want = 0, []
    for a in range(A):
        want.append(max(b5[off:off + n5[a]]))
        off += n5[a]
    assert r["ok"] == [True] * A and r["max_bid"] == want, "genTests-style batch failed its own checks"
    t_dt = torch.tensor([dt5], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t_dt, op=dist.ReduceOp.MAX)
    # config 2 shape in a lock-step batch: CCS22, 20 parties x 32-bit bids per auction
    A2 = 4096   # 512 auctions leave the GPU latency-bound (2,200/s); 2048: 4,800/s, 8192: 7,100/s
    r2 = random.Random(2200 + rank)
    bids2 = [r2.randrange(1 << 31) for _ in range(20 * A2)]
    ev2 = [r2.randrange(20) for _ in range(A2)]
    ids2 = [rank * A2 + a for a in range(A2)]
    eng.ccs22_run(13, [20] * A2, [32] * A2, ev2, bids2, auction_ids=ids2)   # warm-up
    sync()
    t0 = time.perf_counter()
    rr = eng.ccs22_run(13, [20] * A2, [32] * A2, ev2, bids2, auction_ids=ids2)
    sync()
    dt2 = time.perf_counter() - t0
    assert all(rr["max_bid"][20 * a + i] == max(bids2[20 * a:20 * a + 20]) for a in range(A2) for i in range(20)), "CCS22 batch failed its own check"
    t_dt2 = torch.tensor([dt2], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t_dt2, op=dist.ReduceOp.MAX)
    out["config2_ccs22_batch"] = {"auctions_per_s": A2 * world / float(t_dt2.item()), "auctions_per_gpu": A2, "seconds": float(t_dt2.item()),
                                  "sample": f"{A2} independent CCS22 auctions per GPU, 20 parties x 32-bit bids each (the reference: 4.46 s per auction on one core)"}
    out["config5_gentests_batch"] = {"auctions_per_s": A * world / float(t_dt.item()), "auctions_per_gpu": A, "seconds": float(t_dt.item()),
                                     "sample": f"{A} independent auctions per GPU (n ~ U{{1..20}}, c ~ U{{1..32}}), lock-step batch, no exchange"}
    return out


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    pairs = 1500  # per thread per step: ~3000 EC_POINT_mul ~ 2 s of CPU work per step
    for _ in range(args.warmup):
        run_ecmul_ref(cores, 100)
    t_total, mults = 0.0, 0.0
    kind = "reference"
    for s in range(args.steps):
        r = run_ecmul_ref(cores, pairs, seed=1 + s)
        t_total += r["seconds"]
        mults += r["mults"]
        kind = r["kind"]
        cores_used = r["threads"]
    value = mults / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int256 (OpenSSL BN, 64-bit limbs)", "data": "synthetic (seeded PA stream scalars)",
        "config": {"workload": WORKLOAD, "sample": f"{2 * pairs} EC_POINT_mul per thread per step (bounded sample of the 2^21-mult step)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores_used, "kind": kind,
                         "sample": f"{int(mults)} libcrypto EC_POINT_mul (half fixed-base, half variable-base) on {cores_used} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_KIND, help="scalar mults per kind per GPU per step (default 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-seal", action="store_true", help="skip the SEAL auction / proof-verify figures")
    ap.add_argument("--config5-auctions", type=int, default=4096,
                    help="genTests-style auctions per GPU in the config-5 figure (12500 per GPU on 8 GPUs = BASELINE's 10^5)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    pinned_cores = pin_to_gpu_numa_node(local_rank) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pa = importlib.import_module("privacy-auction_b200")
    eng = pa.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    n = args.n

    # ---- synthetic seeded inputs, resident in HBM ---------------------------------
    rng = np.random.default_rng(1234 + rank)
    k_fixed = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8)
    k_base = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8)
    k_var = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8)
    t_kf = torch.from_numpy(k_fixed.copy()).cuda()
    t_kb = torch.from_numpy(k_base.copy()).cuda()
    t_kv = torch.from_numpy(k_var.copy()).cuda()
    t_bases = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
    t_out_f = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
    t_out_v = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    eng.fixed_base_mul_dev(t_kb.data_ptr(), t_bases.data_ptr(), n)  # variable bases P = g^k'
    eng.sync()

    def step_device():
        eng.fixed_base_mul_dev(t_kf.data_ptr(), t_out_f.data_ptr(), n)
        eng.var_base_mul_dev(t_bases.data_ptr(), t_kv.data_ptr(), t_out_v.data_ptr(), n)

    peak = eng.measure_int_peak() if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launches
    eng.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    kstats = eng.profile_end()
    launches = eng.launches - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = 2.0 * n * world * args.steps / (ms_max * 1e-3)

    # ---- end to end through the host-buffer ABI (pinned host memory) -------------
    h_kf = torch.from_numpy(k_fixed.copy()).pin_memory()
    h_kv = torch.from_numpy(k_var.copy()).pin_memory()
    h_bases = t_bases.cpu().pin_memory()
    h_out_f = torch.empty(64 * n, dtype=torch.uint8).pin_memory()
    h_out_v = torch.empty(64 * n, dtype=torch.uint8).pin_memory()

    def step_e2e():
        eng._check(eng.lib.pa_fixed_base_mul(eng.ctx, h_kf.data_ptr(), h_out_f.data_ptr(), n))
        eng._check(eng.lib.pa_var_base_mul(eng.ctx, h_bases.data_ptr(), h_kv.data_ptr(), h_out_v.data_ptr(), n))

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * n * world * e2e_steps / float(e2e_s.item())
    # the e2e result must be the same bytes as the device-resident run
    same = bool(torch.equal(h_out_v.cuda(), t_out_v)) and bool(torch.equal(h_out_f.cuda(), t_out_f))

    # ---- secondary figures of BASELINE.json's metric: proof-verifies/s and auctions/s --------------
    seal = None
    if not args.no_seal:
        seal = seal_figures(eng, pa, rank, world, dist if world > 1 else None, torch, args.config5_auctions)

    if rank == 0:
        var = kstats.get("k_var_base", {"launches": 1, "total_ms": float("nan")})
        fix = kstats.get("k_fixed_base", {"launches": 1, "total_ms": float("nan")})
        var_ms = var["total_ms"] / max(var["launches"], 1)
        fix_ms = fix["total_ms"] / max(fix["launches"], 1)
        sm_max = clocks.get("sm_max_mhz") or 1965.0
        nominal_peak = 64.0 * 148 * sm_max * 1e6
        peak_imad = peak["imad_per_s"]
        peak_wide = peak["imad_wide_per_s"]
        achieved = n * WIDE_VAR / (var_ms * 1e-3)
        nominal = n * FM_VAR * IMAD_PER_FM / (var_ms * 1e-3)
        total_k_ms = sum(v["total_ms"] for v in kstats.values()) or 1.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit integers in 32-bit limbs)", "data": "synthetic (seeded scalars; variable bases g^k' from the fixed-base kernel)",
            "config": {"workload": WORKLOAD, "n_fixed_per_gpu": n, "n_var_per_gpu": n,
                       "l2": "working set per step ~420 MB (scalars, points, Jacobian scratch, outputs) > 126 MB L2; kernels are integer-pipe bound"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 128 * n, "d2h_bytes_per_step": 128 * n,
                    "steps": e2e_steps, "bytes_match_device_run": same, "host_cores_pinned_per_rank": pinned_cores},
            "gpu_launches": int(launches),
            # The path is 256-bit integer arithmetic: neither HBM nor the tensor cores bound it, the
            # integer multiply pipe does ("fmaheavy" in ncu).  achieved = executed 32x32->64 multiply-adds
            # per second, peak = the same instruction in a register-only loop on this GPU.
            "roofline": {"bound": "int-multiply pipe (IMAD.WIDE on fmaheavy); not hbm, not tensor", "kernel": "k_var_base",
                         "achieved": achieved / 1e12, "peak": peak_wide / 1e12, "unit": "T IMAD.WIDE/s", "frac": achieved / peak_wide,
                         "traffic": 3.53e9,  # dram read + write per launch, ncu --set full (profiles/r01f_ncu_full_k_var_base_summary.csv)
                         "traffic_note": "algorithmic bytes are 0.2 GB per launch (64 B point + 32 B scalar in, 96 B Jacobian out); the rest is "
                                         "write-back and refill of per-thread stack lines (window tables and register spills at 80 registers), "
                                         "206 GB/s = 3.1 % of HBM bandwidth, not on the critical path (long_scoreboard 0.48 of 11.6 stall cycles per issue)",
                         "peak_source": "measured on this GPU by pa_measure_int_peak (register-only IMAD.WIDE loop); MEASURED_PEAKS.json has no integer figure",
                         "units_per_launch": n, "per_unit": f"{WIDE_VAR} IMAD.WIDE per variable-base mult (ncu count on this kernel)",
                         "avg_launch_ms": var_ms, "share_of_kernel_time": var["total_ms"] / total_k_ms,
                         "ncu": "fmaheavy pipe 85.1 % active, top stalls math_pipe_throttle and wait; 298 k instructions per multiplication, a third of them IMAD.WIDE",
                         "nominal_algorithm": {"per_unit": f"{FM_VAR} field mults x {IMAD_PER_FM} 32-bit IMAD (SURVEY.md 8d, plain double-and-add)",
                                               "achieved_timad_s": nominal / 1e12, "peak_timad_s": peak_imad / 1e12, "frac": nominal / peak_imad,
                                               "note": "above 1 because GLV + co-Z tables execute ~1,800 field mults instead of 2,900"}},
            "roofline_fixed_base": {"bound": "int-multiply pipe", "kernel": "k_fixed_base", "achieved": n * WIDE_FIXED / (fix_ms * 1e-3) / 1e12,
                                    "peak": peak_wide / 1e12, "unit": "T IMAD.WIDE/s", "frac": n * WIDE_FIXED / (fix_ms * 1e-3) / peak_wide,
                                    "avg_launch_ms": fix_ms, "per_unit": f"{WIDE_FIXED} IMAD.WIDE per fixed-base mult (ncu count)"},
            "roofline_point_stream": point_stream_roofline(kstats, n),
            "kernels": {k: {"launches": v["launches"], "avg_ms": v["total_ms"] / max(v["launches"], 1)} for k, v in kstats.items()},
            "int_peak_measured": peak,
            "rates": {"fixed_base_per_s_per_gpu": n / (fix_ms * 1e-3), "var_base_per_s_per_gpu": n / (var_ms * 1e-3)},
        }
        if seal is not None:
            line["seal"] = seal
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            r = run_ecmul_ref(cores, 20000)  # 40000 EC_POINT_mul per thread, ~12 s on every core
            line["cpu_baseline"] = {"value": r["mults"] / r["seconds"], "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
                                    "sample": f"{int(r['mults'])} libcrypto EC_POINT_mul (half fixed-base, half variable-base) in {r['seconds']:.1f} s"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

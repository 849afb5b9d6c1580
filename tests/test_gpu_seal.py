"""Parity of the CUDA proof / round-logic path, through the C ABI, against (a) the
transcripts of the unmodified reference (tests/golden) and (b) the libcrypto
oracle on seeded random statements, including tampered proofs.  Bit-exact."""
import glob
import os
import random

import pytest

import secp256k1_py as E
import seal_flow

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "seal_*.bin")))
N = E.N


def b32(x):
    return int(x).to_bytes(32, "big")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_engine_reproduces_reference_transcript(engine, path):
    gold = open(path, "rb").read()
    t = seal_flow.parse_transcript(gold)
    fl = seal_flow.SealFlow(engine, t["n"], t["c"], t["seed"], t["bids"])
    out = fl.run()
    assert out == gold
    assert fl.ok and [fl.max_bid] * t["n"] == t["max_bid"]


def test_engine_matches_oracle_on_a_larger_auction(engine, oracle):
    """n = 12, c = 10 (too slow for the all-pairs reference, fine verify-once)"""
    rnd = random.Random(5)
    bids = [rnd.randrange(1 << 10) for _ in range(12)]
    a = seal_flow.SealFlow(engine, 12, 10, 77, bids)
    b = seal_flow.SealFlow(oracle, 12, 10, 77, bids)
    assert a.run() == b.run() and a.ok and b.ok


def _pts(oracle, rnd, k):
    return oracle.fixed_base_mul(b"".join(b32(rnd.randrange(1, N)) for _ in range(k)))


def test_challenge_parity(engine, oracle):
    rnd = random.Random(31)
    for k in (2, 7, 15, 27):
        n = 64
        pts = bytearray(_pts(oracle, rnd, n * k))
        for i in rnd.sample(range(n * k), 10):   # infinity shortens the message
            pts[64 * i:64 * i + 64] = bytes(64)
        ids = [rnd.randrange(1 << 40) for _ in range(n)]
        assert engine.challenge(bytes(pts), k, ids) == oracle.challenge(bytes(pts), k, ids)
    # independent restatement
    pts = _pts(oracle, rnd, 3)
    h = engine.challenge(pts, 3, [9])
    assert int.from_bytes(h, "big") == E.challenge([E.dec64(pts[64 * i:64 * i + 64]) for i in range(3)], 9)


def test_rng_fill_matches_pa_stream(engine):
    streams, ctrs = [0, 1, 5, (3 << 32) | 7], [0, 10, 0, 123456]
    out, ctr2 = engine.rng_fill(42, streams, ctrs, 11)
    for i, (s, c) in enumerate(zip(streams, ctrs)):
        ps = E.PaStream(42, s, c)
        for k in range(11):
            assert out[32 * (11 * i + k):32 * (11 * i + k + 1)] == b32(ps.rand_range())
        assert ctr2[i] == ps.ctr


def test_commit_points_parity(engine, oracle):
    rnd = random.Random(32)
    n = 200
    al = b"".join(b32(rnd.randrange(N)) for _ in range(n - 2)) + b32(0) + b32(N - 1)
    be = b"".join(b32(rnd.randrange(N)) for _ in range(n - 2)) + b32(5) + b32(N - 1)
    bits = bytes(rnd.randrange(2) for _ in range(n))
    assert engine.commit_points(al, be, bits) == oracle.commit_points(al, be, bits)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 20, 127, 128, 129, 1000])
def test_y_scan_and_sum(engine, oracle, n):
    rnd = random.Random(100 + n)
    X = bytearray(_pts(oracle, rnd, n))
    if n >= 5:
        X[64:128] = bytes(64)                      # an infinity among the keys
        X[192:256] = X[128:192]                    # a repeated key (doubling inside the scan)
    X = bytes(X)
    Y = engine.y_scan(X)
    if n <= 129:
        assert Y == oracle.y_scan(X)               # the reference's O(n^2) loops
    else:
        pts = [E.dec64(X[64 * i:64 * i + 64]) for i in range(n)]
        pre = [E.INF]
        for p in pts:
            pre.append(E.add(pre[-1], p))
        tot = pre[-1]
        for i in rnd.sample(range(n), 40) + [0, n - 1]:
            want = E.add(pre[i], E.neg(E.add(tot, E.neg(pre[i + 1]))))
            assert E.dec64(Y[64 * i:64 * i + 64]) == want
    assert engine.point_sum_is_inf(X) == oracle.point_sum_is_inf(X)
    # sum_i Y_i * x_i vanishes when nobody vetoes: use b_i = x_i * Y_i with X_i = x_i * G
    if n <= 129:
        xs = [rnd.randrange(1, N) for _ in range(n)]
        Xs = oracle.fixed_base_mul(b"".join(b32(x) for x in xs))
        Ys = engine.y_scan(Xs)
        bs = engine.var_base_mul(Ys, b"".join(b32(x) for x in xs))
        assert engine.point_sum_is_inf(bs) is True
        assert oracle.point_sum_is_inf(bs) is True


def test_y_scan_batch_matches_single(engine, oracle):
    rnd = random.Random(55)
    sizes = [1, 2, 20, 7, 1, 13, 128, 3]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    X = _pts(oracle, rnd, offs[-1])
    Y = engine.y_scan_batch(X, offs)
    for a, b in zip(offs, offs[1:]):
        assert Y[64 * a:64 * b] == oracle.y_scan(X[64 * a:64 * b])
    flags = engine.point_sum_is_inf_batch(X, offs)
    assert flags == [oracle.point_sum_is_inf(X[64 * a:64 * b]) for a, b in zip(offs, offs[1:])]


def _random_proofs(oracle, rnd, kind, n):
    """valid proofs of every branch, made by the oracle prover from random witnesses"""
    ids = [rnd.randrange(1 << 20) for _ in range(n)]
    sc = lambda: rnd.randrange(1, N)
    G = E.G
    if kind == "pokdlog":
        x = [sc() for _ in range(n)]
        X = oracle.fixed_base_mul(b"".join(map(b32, x)))
        rndb = b"".join(b32(sc()) for _ in range(n))
        return dict(stmt=X, secrets=b"".join(map(b32, x)), ids=ids, rnd=rndb, args=())
    if kind == "powfcom":
        al, be = [sc() for _ in range(n)], [sc() for _ in range(n)]
        bits = bytes(rnd.randrange(2) for _ in range(n))
        stmt = oracle.commit_points(b"".join(map(b32, al)), b"".join(map(b32, be)), bits)
        return dict(stmt=stmt, secrets=b"".join(map(b32, al)), ids=ids, rnd=b"".join(b32(sc()) for _ in range(3 * n)), args=(bits,),
                    wsecrets=b"".join(b32(al[i]) + b32(be[i]) for i in range(n)), wargs=(bits,))
    # stage 1 / stage 2 statements
    x, r, al, be = ([sc() for _ in range(n)] for _ in range(4))
    bits = [rnd.randrange(2) for _ in range(n)]
    X = oracle.fixed_base_mul(b"".join(map(b32, x)))
    R = oracle.fixed_base_mul(b"".join(map(b32, r)))
    Y = oracle.fixed_base_mul(b"".join(b32(sc()) for _ in range(n)))
    cpts = oracle.commit_points(b"".join(map(b32, al)), b"".join(map(b32, be)), bytes(bits))
    base = b"".join((R if bits[i] else Y)[64 * i:64 * i + 64] for i in range(n))
    b = oracle.var_base_mul(base, b"".join(map(b32, x)))
    p = lambda buf, i: buf[64 * i:64 * i + 64]
    if kind == "stage1":
        stmt = b"".join(p(b, i) + p(X, i) + p(Y, i) + p(R, i) + cpts[192 * i:192 * i + 192] for i in range(n))
        sec = b"".join(b32(x[i]) + b32(al[i]) for i in range(n))
        wsec = b"".join(b32(x[i]) + b32(al[i]) + b32(r[i]) + b32(be[i]) for i in range(n))
        return dict(stmt=stmt, secrets=sec, ids=ids, rnd=b"".join(b32(sc()) for _ in range(5 * n)), args=(bytes(bits),),
                    wsecrets=wsec, wargs=(bytes(bits),))
    # stage 2: previous deciding step data; bj random, bi = bit & bj
    xj, rj = [sc() for _ in range(n)], [sc() for _ in range(n)]
    bj = [rnd.randrange(2) for _ in range(n)]
    bi = [bits[i] & bj[i] for i in range(n)]
    Xj = oracle.fixed_base_mul(b"".join(map(b32, xj)))
    Rj = oracle.fixed_base_mul(b"".join(map(b32, rj)))
    Yj = oracle.fixed_base_mul(b"".join(b32(sc()) for _ in range(n)))
    basej = b"".join((Rj if bj[i] else Yj)[64 * i:64 * i + 64] for i in range(n))
    Bj = oracle.var_base_mul(basej, b"".join(map(b32, xj)))
    # the commitment holds the TRUE bit: equal to bi while the bidder is in the race (bj = 1), free once it is out
    cpts = oracle.commit_points(b"".join(map(b32, al)), b"".join(map(b32, be)), bytes(bits))
    basei = b"".join((R if bi[i] else Y)[64 * i:64 * i + 64] for i in range(n))
    Bi = oracle.var_base_mul(basei, b"".join(map(b32, x)))
    stmt = b"".join(p(Bi, i) + p(X, i) + p(R, i) + p(Bj, i) + p(Xj, i) + p(Rj, i) + cpts[192 * i:192 * i + 192] + p(Y, i) + p(Yj, i)
                    for i in range(n))
    sec = b"".join(b32(x[i]) + b32(xj[i]) + b32(al[i]) for i in range(n))
    wsec = b"".join(b32(x[i]) + b32(xj[i]) + b32(al[i]) + b32(r[i]) + b32(rj[i]) + b32(be[i]) for i in range(n))
    return dict(stmt=stmt, secrets=sec, ids=ids, rnd=b"".join(b32(sc()) for _ in range(11 * n)), args=(bytes(bi), bytes(bj)),
                wsecrets=wsec, wargs=(bytes(bi), bytes(bj), bytes(bits)))


REC = {"pokdlog": (96, 1), "powfcom": (352, 4), "stage1": (672, 8), "stage2": (1344, 16)}


@pytest.mark.parametrize("kind", ["pokdlog", "powfcom", "stage1", "stage2"])
def test_prove_verify_parity_and_tampering(engine, oracle, kind):
    rnd = random.Random({"pokdlog": 1, "powfcom": 2, "stage1": 3, "stage2": 4}[kind])
    n = 96
    w = _random_proofs(oracle, rnd, kind, n)
    prove_e, prove_o = getattr(engine, kind + "_prove"), getattr(oracle, kind + "_prove")
    ver_e, ver_o = getattr(engine, kind + "_verify"), getattr(oracle, kind + "_verify")
    proofs = prove_e(w["stmt"], w["secrets"], *w["args"], w["ids"], w["rnd"])
    assert proofs == prove_o(w["stmt"], w["secrets"], *w["args"], w["ids"], w["rnd"])
    if kind != "pokdlog":   # the prover with witnesses (pa_*_prove_w) publishes the same bytes
        assert getattr(engine, kind + "_prove_w")(w["stmt"], w["wsecrets"], *w["wargs"], w["ids"], w["rnd"]) == proofs
    assert ver_e(proofs, w["stmt"], w["ids"]) == bytes([1] * n) == ver_o(proofs, w["stmt"], w["ids"])
    # negative tests (the reference has none): every scalar field flipped,
    # every eps replaced by another valid point, wrong id — verdicts must agree with the oracle
    rec, neps = REC[kind]
    nsc = (rec - 64 * neps) // 32
    bad = bytearray(proofs)
    expect_fail = []
    for i in range(n):
        f = i % (neps + nsc + 1)
        o = rec * i
        if f < neps:
            other = rec * ((i + 1) % n) + 64 * ((f + 1) % neps)
            bad[o + 64 * f:o + 64 * f + 64] = proofs[other:other + 64]
            expect_fail.append(True)
        elif f < neps + nsc:
            pos = o + 64 * neps + 32 * (f - neps) + 31
            bad[pos] ^= 1
            expect_fail.append(True)
        else:
            expect_fail.append(False)
    got = ver_e(bytes(bad), w["stmt"], w["ids"])
    assert got == ver_o(bytes(bad), w["stmt"], w["ids"])
    assert [g == 0 for g in got] == expect_fail
    wrong_ids = [i + 1 for i in w["ids"]]
    assert ver_e(proofs, w["stmt"], wrong_ids) == bytes(n) == ver_o(proofs, w["stmt"], wrong_ids)


# ---- the device-resident whole-auction runner (pa_seal_run) ---------------------------------
@pytest.mark.parametrize("schedule", [1, 2], ids=["step-major", "phase-major"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_runner_reproduces_reference_transcript(engine, path, schedule):
    gold = open(path, "rb").read()
    t = seal_flow.parse_transcript(gold)
    res = engine.seal_run(t["seed"], [t["n"]], [t["c"]], t["bids"], verify=True, sections=True, schedule=schedule)
    assert res["max_bid"] == [max(t["bids"])] and res["ok"] == [True]
    assert seal_flow.sections_to_transcripts(t["seed"], [t["n"]], [t["c"]], t["bids"], res)[0] == gold


@pytest.mark.parametrize("first_one", [0, 1, 2, 3, 4, 7, None], ids=lambda v: f"J={v}")
def test_phase_major_junction_planes_match_oracle(engine, oracle, first_one):
    """The phase-major schedule computes, beside its first window, the steps after the junction for the guesses
    J = 0..3 (pa_seal.cuh, "junction planes") and copies the right plane into place.  Auctions whose first deciding step
    is 0, 1, 2, 3 (a plane is used), 4 and 7 (no plane fits: the windows go on) and never (all-zero bids) must publish the
    oracle's bytes in both schedules."""
    rnd = random.Random(77 + (first_one if first_one is not None else 99))
    n, c = 9, 10
    if first_one is None:
        bids = [0] * n
    else:   # the highest set bit of the maximum is bit c - 1 - first_one
        top = c - first_one
        bids = [rnd.randrange(1 << top) for _ in range(n)]
        bids[rnd.randrange(n)] |= 1 << (top - 1)
    want = seal_flow.SealFlow(oracle, n, c, 123, bids).run()
    for schedule in (2, 1):
        res = engine.seal_run(123, [n], [c], bids, verify=True, sections=True, schedule=schedule)
        assert res["ok"] == [True] and res["max_bid"] == [max(bids)]
        assert seal_flow.sections_to_transcripts(123, [n], [c], bids, res)[0] == want, schedule


def test_runner_batch_of_ragged_auctions_matches_oracle(engine, oracle):
    """genTests-style batch (tests/genTests.py:13-17): n ~ U{1..20}, c ~ U{1..32}, in ONE lock-step run;
    every auction's transcript must equal the oracle's run of that auction alone."""
    rnd = random.Random(404)
    A = 10
    n = [rnd.randint(1, 9) for _ in range(A)]
    c = [rnd.randint(1, 10) for _ in range(A)]
    n[0], c[0], n[1], c[1] = 1, 1, 1, 7           # degenerate shapes
    bids, per = [], []
    for a in range(A):
        b = [rnd.randrange(1 << c[a]) for _ in range(n[a])]
        if a == 2:
            b = [0] * n[a]                         # nobody ever vetoes
        per.append(b)
        bids += b
    ids = [1000 + a for a in range(A)]
    res = engine.seal_run(99, n, c, bids, verify=True, sections=True, auction_ids=ids)
    got = seal_flow.sections_to_transcripts(99, n, c, bids, res)
    assert res["ok"] == [True] * A and res["max_bid"] == [max(b) for b in per]
    for a in range(A):
        want = seal_flow.SealFlow(oracle, n[a], c[a], 99, per[a], auction=ids[a]).run()
        assert got[a] == want, f"auction {a} (n={n[a]}, c={c[a]})"


@pytest.mark.parametrize("schedule", [1, 2], ids=["step-major", "phase-major"])
def test_runner_32_bit_bids(engine, schedule):
    """c = 32 with the top bit set: the reference cannot run this"""
    bids = [0x80000001, 0xFFFFFFFF, 0x7FFFFFFF, 5]
    res = engine.seal_run(3, [4], [32], bids, verify=True, schedule=schedule)
    assert res["ok"] == [True] and res["max_bid"] == [0xFFFFFFFF]


def test_runner_schedules_agree(engine):
    """The phase-major schedule of a single auction publishes the same bytes as the step-major one: junction at
    the first, a middle, the second-to-last and the last step, never (all bids 0), and one bidder alone."""
    rnd = random.Random(88)
    cases = [(7, 9, None), (5, 8, 0x80), (6, 8, 0x02), (6, 8, 0x01), (4, 6, 0), (1, 5, 0x15), (40, 16, None), (3, 1, 1), (3, 2, 2),
             (3, 64, (1 << 64) - 1), (2, 33, 1 << 32)]
    for n, c, top in cases:
        bids = [rnd.randrange(1 << c) for _ in range(n)] if top is None else [rnd.randrange(top + 1) if top else 0 for _ in range(n)]
        if top:
            bids[rnd.randrange(n)] = top
        a = engine.seal_run(500 + n, [n], [c], bids, verify=True, sections=True, schedule=1)
        b = engine.seal_run(500 + n, [n], [c], bids, verify=True, sections=True, schedule=2)
        assert a["ok"] == b["ok"] == [True] and a["max_bid"] == b["max_bid"] == [max(bids)], (n, c, bids)
        ta = seal_flow.sections_to_transcripts(500 + n, [n], [c], bids, a)[0]
        tb = seal_flow.sections_to_transcripts(500 + n, [n], [c], bids, b)[0]
        assert ta == tb, (n, c, bids)


def test_runner_matches_oracle_n64_c12(engine, oracle):
    """a mid-size auction (64 bidders x 12 bits, ~67 k scalar mults on the oracle's single core)"""
    rnd = random.Random(6412)
    bids = [rnd.randrange(1 << 12) for _ in range(64)]
    fl = seal_flow.SealFlow(oracle, 64, 12, 64012, bids)
    want = fl.run()
    for schedule in (1, 2):
        res = engine.seal_run(64012, [64], [12], bids, verify=True, sections=True, schedule=schedule)
        got = seal_flow.sections_to_transcripts(64012, [64], [12], bids, res)[0]
        assert got == want and fl.ok and res["ok"] == [True] and res["max_bid"] == [max(bids)], schedule


def test_runner_all_pairs_work_same_verdicts(engine):
    """verify = n - 1 repeats every verification as often as the reference's n bidders do;
    the published bytes and verdicts are those of verify = 1, in both schedules"""
    rnd = random.Random(31)
    n, c = 6, 7
    bids = [rnd.randrange(1 << c) for _ in range(n)]
    base = engine.seal_run(8, [n], [c], bids, verify=True, sections=True, schedule=1)
    for schedule in (1, 2):
        res = engine.seal_run(8, [n], [c], bids, verify=n - 1, sections=True, schedule=schedule)
        assert res == base

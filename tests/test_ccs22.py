"""CCS22: the oracle port against the transcripts of the
unmodified reference (CPU), and the CUDA engine against both (GPU)."""
import glob
import os
import random
import struct

import pytest

import ccs22_flow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ccs22_*.bin")))


def _hdr(gold):
    n, c, seed, ev = struct.unpack_from("<QQQQ", gold, 8)
    return n, c, seed, ev, list(struct.unpack_from(f"<{n}Q", gold, 40))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_reference_ccs22(oracle, path):
    gold = open(path, "rb").read()
    n, c, seed, ev, bids = _hdr(gold)
    fl = ccs22_flow.Ccs22Flow(oracle, n, c, seed, ev, bids)
    assert fl.run() == gold and fl.max_bid == [max(bids)] * n


def test_setup_hash_zero_scalar_path(oracle):
    import hashlib
    import secp256k1_py as E
    ks = [5, 0x1234, 1 << 200]
    want = int.from_bytes(hashlib.sha256(b"\x05" + b"\x12\x34" + (1 << 200).to_bytes(26, "big")).digest(), "big") % E.N
    assert oracle.ccs22_setup_hash(b"".join(k.to_bytes(32, "big") for k in ks), 3) == want.to_bytes(32, "big")
    assert oracle.ccs22_setup_hash(b"".join(k.to_bytes(32, "big") for k in [5, 0, 7]), 3) == bytes(32)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_engine_reproduces_reference_ccs22(engine, path):
    gold = open(path, "rb").read()
    n, c, seed, ev, bids = _hdr(gold)
    fl = ccs22_flow.Ccs22Flow(engine, n, c, seed, ev, bids)
    assert fl.run() == gold and fl.max_bid == [max(bids)] * n


@pytest.mark.gpu
def test_engine_ccs22_config2_matches_oracle(engine, oracle):
    """BASELINE config 2 shape: 20 bidders, 32-bit bids (the reference itself degenerates to all-zero
    bids at c = 32; here once all-zero and once uniform < 2^31)"""
    rnd = random.Random(22)
    for bids in ([0] * 20, [rnd.randrange(1 << 31) for _ in range(20)]):
        a = ccs22_flow.Ccs22Flow(engine, 20, 32, 5, 7, bids)
        b = ccs22_flow.Ccs22Flow(oracle, 20, 32, 5, 7, bids)
        assert a.run() == b.run() and a.max_bid == [max(bids)] * 20


@pytest.mark.gpu
def test_engine_setup_hash_parity(engine, oracle):
    rnd = random.Random(23)
    for k in (4, 17, 128, 21 * 32):
        sc = bytearray(b"".join(rnd.getrandbits(rnd.choice([256, 250, 200, 64, 8])).to_bytes(32, "big") for _ in range(9 * k)))
        sc[32 * (3 * k + 1):32 * (3 * k + 2)] = bytes(32)     # item 3 has a zero scalar
        assert engine.ccs22_setup_hash(bytes(sc), k) == oracle.ccs22_setup_hash(bytes(sc), k)


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [1, 2], ids=["step-major", "phase-major"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_ccs22_runner_reproduces_reference(engine, path, schedule):
    gold = open(path, "rb").read()
    n, c, seed, ev, bids = _hdr(gold)
    res = engine.ccs22_run(seed, [n], [c], [ev], bids, sections=True, schedule=schedule)
    assert ccs22_flow.sections_to_transcripts(seed, [n], [c], [ev], bids, res)[0] == gold
    assert res["max_bid"] == [max(bids)] * n


@pytest.mark.gpu
def test_ccs22_runner_ragged_batch_matches_oracle(engine, oracle):
    rnd = random.Random(909)
    A = 8
    n = [rnd.randint(1, 7) for _ in range(A)]
    c = [rnd.randint(1, 9) for _ in range(A)]
    n[0], c[0] = 1, 2
    ev = [rnd.randrange(x) for x in n]
    per = [[rnd.randrange(1 << c[a]) for _ in range(n[a])] for a in range(A)]
    per[1] = [0] * n[1]
    bids = [b for row in per for b in row]
    ids = [500 + a for a in range(A)]
    res = engine.ccs22_run(31, n, c, ev, bids, sections=True, auction_ids=ids)
    got = ccs22_flow.sections_to_transcripts(31, n, c, ev, bids, res)
    for a in range(A):
        fl = ccs22_flow.Ccs22Flow(oracle, n[a], c[a], 31, ev[a], per[a], auction=ids[a])
        assert got[a] == fl.run(), f"auction {a} (n={n[a]}, c={c[a]}, evaluator={ev[a]})"


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [1, 2], ids=["step-major", "phase-major"])
@pytest.mark.parametrize("name", ["ccs22_20_32", "ccs22_20_31"])
def test_ccs22_runner_config2_digest(engine, name, schedule):
    import hashlib, json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_config_digests.json")))[name]
    res = engine.ccs22_run(g["seed"], [g["n"]], [g["c"]], [g["evaluator"]], g["bids"], sections=True, schedule=schedule)
    out = ccs22_flow.sections_to_transcripts(g["seed"], [g["n"]], [g["c"]], [g["evaluator"]], g["bids"], res)[0]
    assert hashlib.sha256(out).hexdigest() == g["sha256"]


@pytest.mark.gpu
def test_ccs22_runner_schedules_agree(engine):
    """phase-major == step-major on shapes the goldens do not have: 40 parties, evaluator first / last, the evaluator
    holding the maximum (alpha = 1 steps), everybody bidding 0, and one party alone"""
    rnd = random.Random(4040)
    for n, c, ev, kind in [(40, 12, 0, "rand"), (9, 10, 8, "evmax"), (6, 7, 2, "zero"), (1, 4, 0, "rand"), (2, 1, 1, "rand"), (12, 32, 5, "rand"), (3, 64, 1, "evmax")]:
        bids = [0] * n if kind == "zero" else [rnd.randrange(1 << (c - 1)) for _ in range(n)]
        if kind == "evmax":
            bids[ev] = (1 << c) - 1
        a = engine.ccs22_run(77, [n], [c], [ev], bids, sections=True, schedule=1)
        b = engine.ccs22_run(77, [n], [c], [ev], bids, sections=True, schedule=2)
        assert a["max_bid"] == b["max_bid"] == [max(bids)] * n, (n, c, ev, kind)
        for key in a:
            assert a[key] == b[key], (n, c, ev, kind, key)


@pytest.mark.gpu
def test_fused_ot_messages_match_oracle_composition(engine, oracle):
    """pa_ccs22_ot_recv1 / pa_ccs22_ot_send (one warp per scalar multiplication) against the same
    messages composed from the oracle's EC_POINT_mul call shapes (CCS22/evaluator.cpp:91-111, bidder.cpp:155-198)"""
    import secp256k1_py as E
    rnd = random.Random(515)
    n = 70
    sc = lambda: b"".join(rnd.getrandbits(256).to_bytes(32, "big") for _ in range(n))
    k, beta, s, t, m = sc(), sc(), sc(), sc(), sc()
    alpha = b"".join((rnd.randrange(2)).to_bytes(32, "big") for _ in range(n))
    gh = oracle.fixed_base_mul(sc() + sc())
    params = b"".join(gh[64 * i:64 * i + 64] + gh[64 * (n + i):64 * (n + i) + 64] for i in range(n))
    g1 = b"".join(params[128 * i:128 * i + 64] for i in range(n))
    h = b"".join(params[128 * i + 64:128 * i + 128] for i in range(n))
    B = bytearray(oracle.fixed_base_mul(sc()))
    B[64 * 5:64 * 6] = bytes(64)                                  # B = infinity (n = 1 auctions)
    B = bytes(B)
    T2 = oracle.fixed_base_mul(k)
    G = oracle.double_mul(beta, g1, alpha)
    H = oracle.lincomb2(T2, alpha, h, beta)
    want_r1 = b"".join(T2[64 * i:64 * i + 64] + G[64 * i:64 * i + 64] + H[64 * i:64 * i + 64] for i in range(n))
    assert engine.ccs22_ot_recv1(k, beta, alpha, params) == want_r1
    M1 = oracle.fixed_base_mul(m)
    z = oracle.double_mul(s, h, t)
    C0 = oracle.point_add(oracle.lincomb2(G, s, H, t), B)
    C1 = oracle.point_add(oracle.lincomb2(oracle.point_add(G, g1, sub=True), s, oracle.point_add(H, T2, sub=True), t), M1)
    want_s = b"".join(z[64 * i:64 * i + 64] + C0[64 * i:64 * i + 64] + C1[64 * i:64 * i + 64] for i in range(n))
    st = b"".join(s[32 * i:32 * i + 32] + t[32 * i:32 * i + 32] for i in range(n))
    assert engine.ccs22_ot_send(want_r1, params, B, st, m) == want_s


@pytest.mark.gpu
def test_commit_bes_encode_recv2_match_oracle_composition(engine, oracle):
    """pa_ccs22_commit / pa_ccs22_bes_encode / pa_ccs22_ot_recv2 against the same steps composed from the
    oracle's call shapes (CCS22/bidder.cpp:80-88, 118-147; evaluator.cpp:117-156)"""
    import secp256k1_py as E
    rnd = random.Random(616)
    sc = lambda n: b"".join(rnd.getrandbits(256).to_bytes(32, "big") for _ in range(n))
    cut = lambda b, i, w=64: b[w * i:w * i + w]
    # commit: 5 parties, 12 hashed scalars each, one of them a bid of 0
    n, k = 5, 12
    hashed, R = sc(n * k), sc(n)
    bids = [0, 1, 77, 2**31 - 1, 2**32 - 1]
    bid = b"".join(v.to_bytes(32, "big") for v in bids)
    gh = oracle.fixed_base_mul(sc(2 * n))
    params = b"".join(cut(gh, i) + cut(gh, n + i) for i in range(n))
    g1 = b"".join(cut(params, 2 * i) for i in range(n))
    h = b"".join(cut(params, 2 * i + 1) for i in range(n))
    H = oracle.ccs22_setup_hash(hashed, k)
    com = oracle.point_add(oracle.double_mul(bid, g1, H), oracle.var_base_mul(h, R))
    assert engine.ccs22_commit(hashed, k, bid, R, params) == (H, com)
    # BESEncode: 9 public keys, every party once vetoing and once not, plus n = 1 (Y = infinity)
    npk = 9
    X = oracle.fixed_base_mul(sc(npk))
    Y = oracle.y_scan(X)
    ids = list(range(npk)) * 2
    d = [0] * npk + [1] * npk
    x, r = sc(2 * npk), sc(2 * npk)
    want = b"".join(cut(oracle.fixed_base_mul(cut(r, i, 32)), 0) if d[i] else cut(oracle.var_base_mul(cut(Y, ids[i]), cut(x, i, 32)), 0)
                    for i in range(2 * npk))
    assert engine.ccs22_bes_encode(X, ids, d, x, r) == want
    assert engine.ccs22_bes_encode(cut(X, 0), [0], [0], cut(x, 0, 32), cut(r, 0, 32)) == bytes(64)
    # OTReceive2: sum of C0_j - beta_j z_j plus B; made to cancel exactly, then perturbed
    m = 7
    beta, zk, ck = sc(m), sc(m), sc(m)
    z, C0 = oracle.fixed_base_mul(zk), oracle.fixed_base_mul(ck)
    terms = oracle.point_add(C0, oracle.var_base_mul(z, beta), sub=True)
    total = E.INF
    for j in range(m):
        total = E.add(total, E.dec64(cut(terms, j)))
    ots = b"".join(cut(z, j) + cut(C0, j) + cut(C0, j) for j in range(m))
    assert engine.ccs22_ot_recv2(ots, beta, E.enc64(E.neg(total))) is True
    assert engine.ccs22_ot_recv2(ots, beta, E.enc64(total)) is False
    assert engine.ccs22_ot_recv2(b"", b"", bytes(64)) is True          # n = 1 auction: no other party, B = infinity
    assert engine.ccs22_ot_recv2(b"", b"", cut(z, 0)) is False

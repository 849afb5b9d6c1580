"""Parity of the CUDA scalar-multiplication path against the oracle, through the
C ABI, on seeded inputs — bit-exact (integer work)."""
import random

import pytest

import secp256k1_py as E

pytestmark = pytest.mark.gpu
N, M = E.N, 2**256


def _b(ks):
    return b"".join(k.to_bytes(32, "big") for k in ks)


EDGE_K = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 255, 256, N - 1, N - 2, N, N + 5, M - 1, (N - 1) // 2,
          int("8" * 64, 16) % N, int("7" * 64, 16), 2**255, 2**128, 1 << 8, 1 << 248, 0xFF << 248]
# boundaries of the signed 16-bit fixed-base windows: digits -2^15, 2^15 - 1, carries rippling through every window
EDGE_K += [int(w * (64 // len(w)), 16) for w in ("8000", "7FFF", "8001", "0001", "FFFF", "7FFF8000", "80007FFF", "0000FFFF")] + \
          [(1 << 15) - 1, 1 << 15, (1 << 15) + 1, (1 << 16) - 1, 1 << 16, 0x7FFF << 240, 0x8000 << 240, (1 << 256) - (1 << 16)]


def _scalars(seed, n):
    rnd = random.Random(seed)
    return EDGE_K + [rnd.getrandbits(256) for _ in range(n - len(EDGE_K))]


def test_fixed_base_parity(engine, oracle):
    ks = _scalars(21, 4096)
    got = engine.fixed_base_mul(_b(ks))
    assert got == oracle.fixed_base_mul(_b(ks))
    # independent restatement on a few
    for i in list(range(len(EDGE_K))) + [100, 4095]:
        assert E.dec64(got[64 * i:64 * i + 64]) == E.mul(ks[i], E.G)


def test_var_base_parity(engine, oracle):
    ks = _scalars(22, 2048)
    base_k = _scalars(23, 2048)[::-1]          # includes 0 (infinity base) and n (infinity)
    pts = oracle.fixed_base_mul(_b(base_k))
    assert engine.var_base_mul(pts, _b(ks)) == oracle.var_base_mul(pts, _b(ks))


def test_var_base_edge_scalars_on_one_point(engine, oracle):
    p = oracle.fixed_base_mul(_b([0xC0FFEE]))
    ks = EDGE_K
    pts = p * len(ks)
    assert engine.var_base_mul(pts, _b(ks)) == oracle.var_base_mul(pts, _b(ks))


def test_double_mul_parity(engine, oracle):
    a, b = _scalars(24, 2048), _scalars(25, 2048)[::-1]
    pts = oracle.fixed_base_mul(_b(_scalars(26, 2048)))
    assert engine.double_mul(_b(a), pts, _b(b)) == oracle.double_mul(_b(a), pts, _b(b))


def test_double_mul_cancellation(engine, oracle):
    """a*G + b*(k*G) hitting infinity and the doubling case inside the final add."""
    k = 0xABCDEF
    kinv = pow(k, -1, N)
    p = oracle.fixed_base_mul(_b([k]))
    a = [5, N - 5, 7, 0, 9]
    b = [(N - 5) * kinv % N, 5 * kinv % N, 7 * kinv % N, 0, 0]
    pts = p * len(a)
    got = engine.double_mul(_b(a), pts, _b(b))
    assert got == oracle.double_mul(_b(a), pts, _b(b))
    assert got[:128] == bytes(128)  # the first two cancel to infinity


def test_lincomb2_parity(engine, oracle):
    a, b = _scalars(27, 1024), _scalars(28, 1024)[::-1]
    p = oracle.fixed_base_mul(_b(_scalars(29, 1024)))
    q = oracle.fixed_base_mul(_b(_scalars(30, 1024)[::-1]))
    assert engine.lincomb2(p, _b(a), q, _b(b)) == oracle.lincomb2(p, _b(a), q, _b(b))
    # same base twice, and base + its negative
    assert engine.lincomb2(p, _b(a), p, _b(b)) == oracle.lincomb2(p, _b(a), p, _b(b))
    negp = oracle.point_add(bytes(len(p)), p, sub=True)
    assert engine.lincomb2(p, _b(a), negp, _b(a)) == bytes(len(p))


def test_point_add_sub_encode(engine, oracle):
    p = oracle.fixed_base_mul(_b(_scalars(31, 512)))
    q = oracle.fixed_base_mul(_b(_scalars(32, 512)[::-1]))
    assert engine.point_add(p, q) == oracle.point_add(p, q)
    assert engine.point_add(p, q, sub=True) == oracle.point_add(p, q, sub=True)
    assert engine.point_add(p, p) == oracle.point_add(p, p)
    assert engine.point_add(p, p, sub=True) == bytes(len(p))
    assert engine.point_encode(p) == oracle.point_encode(p)
    assert engine.point_encode(p, compressed=True) == oracle.point_encode(p, compressed=True)


def test_empty_and_ragged_batches(engine, oracle):
    assert engine.fixed_base_mul(b"") == b""
    for n in (1, 31, 33, 127, 129, 1000):
        ks = _scalars(40 + n, max(n, len(EDGE_K)))[:n]
        assert engine.fixed_base_mul(_b(ks)) == oracle.fixed_base_mul(_b(ks))


def test_full_size_properties(engine, oracle):
    """2^20 scalars (BASELINE config 3): linearity k*G + k'*G == (k+k')*G through
    three different code paths, plus a seeded sample against libcrypto."""
    n = 1 << 20
    rnd = random.Random(99)
    k1 = [rnd.getrandbits(256) for _ in range(n)]
    k2 = [rnd.getrandbits(256) for _ in range(n)]
    g1 = engine.fixed_base_mul(_b(k1))
    g2 = engine.fixed_base_mul(_b(k2))
    gs = engine.fixed_base_mul(_b([(a + b) % N for a, b in zip(k1, k2)]))
    assert engine.point_add(g1, g2) == gs
    # variable base: k2 * (k1 * G) == (k1 * k2) * G
    v = engine.var_base_mul(g1, _b(k2))
    assert v == engine.fixed_base_mul(_b([(a % N) * (b % N) % N for a, b in zip(k1, k2)]))
    idx = rnd.sample(range(n), 512)
    sub_k = _b([k1[i] for i in idx])
    assert b"".join(g1[64 * i:64 * i + 64] for i in idx) == oracle.fixed_base_mul(sub_k)
    sub_p = b"".join(g1[64 * i:64 * i + 64] for i in idx)
    assert b"".join(v[64 * i:64 * i + 64] for i in idx) == oracle.var_base_mul(sub_p, _b([k2[i] for i in idx]))


def test_point_on_curve(engine, oracle):
    p = bytearray(oracle.fixed_base_mul(_b(_scalars(61, 64))))
    want = [1] * 64
    p[64 * 3 + 63] ^= 1; want[3] = 0                                   # y off by one
    p[64 * 5:64 * 5 + 32] = (E.P + 1).to_bytes(32, "big"); want[5] = 0   # x >= p
    p[64 * 7:64 * 8] = bytes(64); want[7] = 1                            # infinity
    x0 = int.from_bytes(p[64 * 9:64 * 9 + 32], "big")
    p[64 * 9 + 32:64 * 10] = ((E.P - int.from_bytes(p[64 * 9 + 32:64 * 10], "big")) % E.P).to_bytes(32, "big")  # -P is on the curve
    got = engine.point_on_curve(bytes(p))
    assert list(got) == want
    for i in range(64):
        assert bool(got[i]) == E.on_curve(E.dec64(bytes(p[64 * i:64 * i + 64]))) or i == 5


def test_pipeline_chunk_boundaries(engine, oracle):
    """Host-buffer calls above one pipeline chunk (148 x 128 x 20 items) are cut into a short head, whole
    waves and a short tail; sizes around every boundary must give the same bytes as the oracle on a
    sample and as the one-chunk path everywhere."""
    unit = 148 * 128
    rnd = random.Random(7)
    base = [rnd.getrandbits(256) for _ in range(4096)]
    for n in (20 * unit, 20 * unit + 1, 25 * unit - 1, 25 * unit + 3, 45 * unit + 77):
        ks = (base * (n // len(base) + 1))[:n]
        got = engine.fixed_base_mul(_b(ks))
        one = engine.fixed_base_mul(_b(base))                     # single-chunk path
        assert len(got) == 64 * n
        for off in range(0, n, len(base)):
            m = min(len(base), n - off)
            assert got[64 * off:64 * (off + m)] == one[:64 * m], (n, off)
    assert one[:64 * 64] == oracle.fixed_base_mul(_b(base[:64]))


def test_two_contexts_and_argument_errors(pa, oracle):
    """Contexts are independent (one per thread is the rule): two alive at once, calls interleaved.
    Bad arguments come back as an error code with a message, not a crash."""
    a, b = pa.Engine(0), pa.Engine(0)
    try:
        ks = _scalars(77, 300)
        want = oracle.fixed_base_mul(_b(ks))
        ra = a.fixed_base_mul(_b(ks))
        pts = b.fixed_base_mul(_b(ks[::-1]))
        assert ra == want and a.var_base_mul(pts, _b(ks)) == b.var_base_mul(pts, _b(ks)) == oracle.var_base_mul(pts, _b(ks))
        import ctypes
        rc = a.lib.pa_fixed_base_mul(a.ctx, None, None, 5)
        assert rc < 0 and a.lib.pa_last_error(a.ctx)
        out = ctypes.create_string_buffer(64)
        off_curve = (5).to_bytes(32, "big") + (7).to_bytes(32, "big")
        assert a.point_on_curve(off_curve + pts[:64]) == bytes([0, 1])
        assert a.fixed_base_mul(_b(ks)) == want                     # still usable after an error
    finally:
        a.close()
        b.close()


def test_scalar_mul_jobs_match_single_calls(engine, oracle):
    """pa_scalar_mul_jobs: four job kinds of different lengths in one interleaved pipeline (one of them longer than a
    pipeline chunk, so that chunks of different jobs really alternate) == the single calls == the oracle on a sample."""
    import importlib
    E_ = importlib.import_module("privacy-auction_b200.engine")
    rnd = random.Random(909)
    sizes = {"fixed": 400000, "var": 3000, "double": 1500, "lincomb2": 700}
    sc = lambda n: b"".join(rnd.getrandbits(256).to_bytes(32, "big") for _ in range(n))
    kf = sc(sizes["fixed"])
    pts = engine.fixed_base_mul(sc(sizes["var"]))
    kv, ka, kb = sc(sizes["var"]), sc(sizes["double"]), sc(sizes["double"])
    la, lb = sc(sizes["lincomb2"]), sc(sizes["lincomb2"])
    jobs = [(E_.MUL_FIXED, dict(a=kf), sizes["fixed"]),
            (E_.MUL_VAR, dict(p=pts, a=kv), sizes["var"]),
            (E_.MUL_DOUBLE, dict(a=ka, p=pts[:64 * sizes["double"]], b=kb), sizes["double"]),
            (E_.MUL_LINCOMB2, dict(p=pts[:64 * sizes["lincomb2"]], a=la, q=pts[64 * 100:64 * (100 + sizes["lincomb2"])], b=lb), sizes["lincomb2"])]
    out = engine.scalar_mul_jobs(jobs)
    assert out[0] == engine.fixed_base_mul(kf)
    assert out[1] == engine.var_base_mul(pts, kv) == oracle.var_base_mul(pts, kv)
    assert out[2] == engine.double_mul(ka, pts[:64 * sizes["double"]], kb)
    assert out[3] == engine.lincomb2(pts[:64 * sizes["lincomb2"]], la, pts[64 * 100:64 * (100 + sizes["lincomb2"])], lb)
    assert out[0][:64 * 500] == oracle.fixed_base_mul(kf[:32 * 500])
    assert engine.scalar_mul_jobs([]) == []

"""Limb-level check of the device arithmetic headers (pa_fe.cuh, pa_sc.cuh,
pa_ec.cuh, pa_smul.cuh) compiled by g++ with the PTX carry flag emulated
(tests/hostcheck/hostcheck.cpp) against Python integers.  This is a unit test of
the algorithms; the product library contains device code only."""
import ctypes
import os
import random
import subprocess

import pytest

import secp256k1_py as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "tests", "hostcheck")
P, N, M = E.P, E.N, 2**256


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC, "libhostcheck.so")
    subprocess.run(["g++", "-O2", "-x", "c++", "-std=c++17", "-fPIC", "-shared",
                    "-I" + os.path.join(ROOT, "privacy-auction_b200", "csrc"), "-o", so,
                    os.path.join(HC, "hostcheck.cpp")], check=True)
    return ctypes.CDLL(so)


def b32(x):
    return x.to_bytes(32, "big")


EDGE = [0, 1, 2, P - 1, P, P + 1, M - 1, M - 2, P - 2, 977, 2**32, 2**32 + 977, M - 2**32, 2**255, 2**128 - 1,
        M - 2**224] + [(M - 1) & ~((1 << (32 * i + 32)) - (1 << (32 * i))) for i in range(8)] + \
       [((1 << (32 * i + 32)) - (1 << (32 * i))) for i in range(8)]


def test_field_ops(hc):
    rnd = random.Random(1)
    vals = EDGE + [rnd.getrandbits(256) for _ in range(120)]

    def op(o, a, b=0):
        out = ctypes.create_string_buffer(32)
        hc.hc_fe_op(o, b32(a), b32(b), out)
        return int.from_bytes(out.raw, "big")

    def mp(a, b, sqr=0):
        out = ctypes.create_string_buffer(64)
        hc.hc_mp_mul8(b32(a), b32(b), out, sqr)
        return int.from_bytes(out.raw, "big")

    for a in vals:
        for b in rnd.sample(vals, 12) + EDGE[:16]:
            assert mp(a, b) == a * b
            assert op(0, a, b) == a * b % P
            assert op(2, a, b) == (a + b) % P
            assert op(3, a, b) == (a - b) % P
            assert op(13, a, b) == (a - 2 * b) % P
            assert op(14, a, b) == (-2 * b) % P
            assert op(15, a, b) == (-a - 3 * b) % P
            assert op(8, a, b) == int((a - b) % P == 0)
        assert mp(a, a, 1) == a * a
        assert op(1, a) == a * a % P
        assert op(5, a) == (-a) % P
        for k in (1, 2, 3):
            assert op(8 + k, a) == (a << k) % P
        assert op(12, a) == 3 * a % P
        assert op(6, a) == a % P
        assert op(7, a) == int(a % P == 0)
    for a in vals[:40]:
        if a % P:
            assert op(4, a) == pow(a, P - 2, P)


def test_scalar_ops(hc):
    rnd = random.Random(2)
    vals = [0, 1, 2, N - 1, N - 2, N, N + 1, M - 1, 2**255, 2**128, N // 2] + [rnd.getrandbits(256) for _ in range(100)]

    def op(o, a, b=0):
        out = ctypes.create_string_buffer(32)
        hc.hc_sc_op(o, b32(a), b32(b), out)
        return int.from_bytes(out.raw, "big")

    for a in vals:
        for b in rnd.sample(vals, 10):
            ar, br = a % N, b % N
            assert op(0, a, b) == ar * br % N
            assert op(1, a, b) == (ar + br) % N
            assert op(2, a, b) == (ar - br) % N
        assert op(3, a) == (-a) % N


def test_group_and_scalar_mult(hc):
    rnd = random.Random(3)

    def fixed(k):
        out = ctypes.create_string_buffer(64)
        hc.hc_fixed_base(b32(k), out)
        return E.dec64(out.raw)

    def var(p, k):
        out = ctypes.create_string_buffer(64)
        hc.hc_var_base(E.enc64(p), b32(k), out)
        return E.dec64(out.raw)

    def lin(p, a, q, b):
        out = ctypes.create_string_buffer(64)
        hc.hc_lincomb2(E.enc64(p), b32(a), E.enc64(q), b32(b), out)
        return E.dec64(out.raw)

    def padd(p, q, m):
        out = ctypes.create_string_buffer(64)
        hc.hc_point_add(E.enc64(p), E.enc64(q), out, m)
        return E.dec64(out.raw)

    ks = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 255, 256, N - 1, N - 2, N, N + 5, M - 1, (N - 1) // 2,
          int("8" * 64, 16) % N, int("7" * 64, 16)] + [rnd.getrandbits(256) for _ in range(12)]
    # boundaries of the signed 16-bit fixed-base windows (digits -2^15 and 2^15 - 1, carries through every window)
    wk = [int(w * (64 // len(w)), 16) for w in ("8000", "7FFF", "8001", "0001", "FFFF", "7FFF8000", "80007FFF", "0000FFFF")] + \
         [(1 << 15) - 1, 1 << 15, (1 << 15) + 1, (1 << 16) - 1, 1 << 16, 0x7FFF << 240, 0x8000 << 240, (1 << 256) - (1 << 16)]
    for k in wk:
        assert fixed(k) == E.mul(k, E.G)
    for k in ks:
        assert fixed(k) == E.mul(k, E.G)
    pts = [E.G, E.mul(2, E.G), E.mul(12345, E.G), E.neg(E.G), E.INF, E.mul(N - 1, E.G)] + \
          [E.mul(rnd.getrandbits(256), E.G) for _ in range(3)]
    for p in pts:
        for q in pts:
            for m in (0, 1):
                assert padd(p, q, m) == E.add(p, q)
            assert hc.hc_jac_eq_aff(E.enc64(p), E.enc64(q)) == int(p == q)
    for p in pts[:6]:
        for k in ks[:20] + ks[-3:]:
            assert var(p, k) == E.mul(k, p)
    for _ in range(12):
        p, q, a, b = rnd.choice(pts), rnd.choice(pts), rnd.choice(ks), rnd.choice(ks)
        assert lin(p, a, q, b) == E.lincomb(a, p, b, q)
    p = pts[2]
    for a, b in [(5, N - 5), (1, 1), (0, 0), (7, 0), (0, 9), (N - 1, 1), (8, 8)]:
        assert lin(p, a, p, b) == E.mul(a + b, p)
        assert lin(p, a, E.neg(p), b) == E.mul(a - b, p)


def test_glv_halves_stay_below_2_128(hc):
    """The variable-base loop walks 32 signed 4-bit windows plus a 0/1 top digit (pa_smul.cuh, PA_GLV_WINDOWS): that needs
    |k1|, |k2| < 2^128 for every scalar.  Random scalars, the edges, and scalars next to the rounding boundaries of the
    two quotient estimates c_i = round(k g_i / 2^384)."""
    LAM = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
    G1 = 0x3086D221A7D46BCDE86C90E49284EB153DAA8A1471E8CA7FE893209A45DBB031
    G2 = 0xE4437ED6010E88286F547FA90ABFE4C4221208AC9DF506C61571B4AE8AC47F71
    rnd = random.Random(2026)
    ks = [0, 1, 2, N - 1, N - 2, N // 2, N // 2 + 1, N // 3, LAM, N - LAM] + [rnd.randrange(N) for _ in range(3000)]
    for j in range(1, 400):
        for g in (G1, G2):
            k = (((2 * j + 1) << 383) // g) % N
            ks += [(k - 1) % N, k, (k + 1) % N]
    out = ctypes.create_string_buffer(42)
    worst = 0
    for k in ks:
        hc.hc_glv_split(b32(k), out)
        k1, k2 = int.from_bytes(out.raw[:20], "little"), int.from_bytes(out.raw[20:40], "little")
        worst = max(worst, k1, k2)
        s1, s2 = (-k1 if out.raw[40] else k1), (-k2 if out.raw[41] else k2)
        assert (s1 + s2 * LAM - k) % N == 0
    assert worst < 1 << 128


def test_challenge_hash_and_draw_stream_all_alignments(hc):
    """pa_sha256.cuh streams a message through a 64-bit shift register, whole words for the coordinates of a point, single
    bytes for the 04 / 00 prefixes and the id: every number of points (every alignment of the block boundaries), with
    points at infinity (one byte instead of 65) at every position of a short message and at random positions of long ones,
    against hashlib; and the PA draw stream against its definition."""
    rnd = random.Random(4242)
    pts = [E.mul(rnd.randrange(1, N), E.G) for _ in range(8)]

    def chal(points, ident):
        buf = b"".join(E.enc64(p) for p in points)
        out = ctypes.create_string_buffer(32)
        hc.hc_challenge(buf, len(points), ctypes.c_ulonglong(ident), out)
        return int.from_bytes(out.raw, "big")

    for k in range(0, 29):
        ps = [rnd.choice(pts) for _ in range(k)]
        ident = rnd.getrandbits(64)
        assert chal(ps, ident) == E.challenge(ps, ident), k
        for _ in range(3):   # some of them at infinity
            qs = [E.INF if rnd.random() < 0.3 else p for p in ps]
            assert chal(qs, ident) == E.challenge(qs, ident), k
    for k in range(1, 6):      # infinity at every position
        for pos in range(k):
            qs = [E.INF if i == pos else pts[i] for i in range(k)]
            assert chal(qs, 7) == E.challenge(qs, 7)
    assert chal([E.INF] * 27, 1) == E.challenge([E.INF] * 27, 1)
    out = ctypes.create_string_buffer(32)
    for seed, stream, ctr in [(0, 0, 0), (42, 5, 7), (2**64 - 1, 2**64 - 1, 2**64 - 1), (1, (3 << 32) | 9, 123456789)]:
        hc.hc_stream_draw(ctypes.c_ulonglong(seed), ctypes.c_ulonglong(stream), ctypes.c_ulonglong(ctr), out)
        assert int.from_bytes(out.raw, "big") == E.pa_draw(seed, stream, ctr)


"""The libcrypto-based checker (oracle/pa_oracle.c) against the independent
pure-Python restatement (oracle/secp256k1_py.py) and against libcrypto's own
curve constants."""
import random

import secp256k1_py as E


def _ks(rnd, n):
    edge = [0, 1, 2, E.N - 1, E.N, E.N + 3, 2**256 - 1, 2**255]
    return edge + [rnd.getrandbits(256) for _ in range(n - len(edge))]


def _b(ks):
    return b"".join(k.to_bytes(32, "big") for k in ks)


def _pts(b):
    return [E.dec64(b[i:i + 64]) for i in range(0, len(b), 64)]


def test_curve_constants_from_libcrypto(oracle):
    assert oracle.curve_constants() == (E.P, E.N, E.GX, E.GY)
    assert E.on_curve(E.G) and E.mul(E.N, E.G) is E.INF


def test_scalar_mult_shapes(oracle):
    rnd = random.Random(11)
    ks, ks2 = _ks(rnd, 24), _ks(rnd, 24)[::-1]
    pts_b = oracle.fixed_base_mul(_b(ks))
    pts = _pts(pts_b)
    assert pts == [E.mul(k, E.G) for k in ks]
    assert _pts(oracle.var_base_mul(pts_b, _b(ks2))) == [E.mul(k, p) for k, p in zip(ks2, pts)]
    assert _pts(oracle.double_mul(_b(ks2), pts_b, _b(ks))) == [E.lincomb(a, E.G, b, p) for a, b, p in zip(ks2, ks, pts)]
    q_b = oracle.var_base_mul(pts_b, _b(ks2))
    assert _pts(oracle.lincomb2(pts_b, _b(ks), q_b, _b(ks2))) == [
        E.lincomb(a, p, b, q) for a, p, b, q in zip(ks, pts, ks2, _pts(q_b))]


def test_add_sub_encode(oracle):
    rnd = random.Random(12)
    ks = _ks(rnd, 16)
    p_b = oracle.fixed_base_mul(_b(ks))
    q_b = oracle.fixed_base_mul(_b(ks[::-1]))
    p, q = _pts(p_b), _pts(q_b)
    assert _pts(oracle.point_add(p_b, q_b)) == [E.add(a, b) for a, b in zip(p, q)]
    assert _pts(oracle.point_add(p_b, q_b, sub=True)) == [E.add(a, E.neg(b)) for a, b in zip(p, q)]
    assert _pts(oracle.point_add(p_b, p_b)) == [E.add(a, a) for a in p]
    assert _pts(oracle.point_add(p_b, p_b, sub=True)) == [E.INF] * len(p)
    assert oracle.point_encode(p_b) == [E.point2oct(a) for a in p]
    assert oracle.point_encode(p_b, compressed=True) == [E.point2oct(a, True) for a in p]
